"""BASELINE.json configs at their stated size on one B200 (log: profiles/r02_configs.log).

  python scripts/configs_at_size.py toric15 [p]      # configs[2]: Toric L = 15 through CssCode, 15 rounds
  python scripts/configs_at_size.py unrot15 [p]      # configs[2]: Unrotated d = 15, 15 rounds
  python scripts/configs_at_size.py steane [shots]   # configs[3]: Steane CssCode, low-p tail, 10^9 shots per point
  python scripts/configs_at_size.py bb144 [shots]    # configs[4]: BB [[144,12,12]], 12 rounds, BP + OSD (run under torchrun for N GPUs)

For the matching configs: sizes asserted against SURVEY 8a, the matching graph compared edge by edge with the oracle's
own build from the DEM, a sample of shots decoded with weights / corrections and checked against the exact-MWPM oracle
(built from the DEM, not from the GPU's edge list), then throughput through gq_collect.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

import graphqec_b200 as gq  # noqa: E402
from graphqec_b200 import backend as be  # noqa: E402
from make_reference_circuits import BB144, H_STEANE, toric_checks  # noqa: E402


def b8(t, nbits):
    a = t.cpu().numpy().view(np.uint8).reshape(t.shape[0], -1)
    return np.ascontiguousarray(a[:, : max(1, (nbits + 7) // 8)])


def matching_config(name, code, want_D, want_M_plain, check_shots, rate_shots):
    from oracle import pyoracle as po

    t0 = time.perf_counter()
    code.build_memory_circuit(number_of_rounds=code.distance, logic_check="Z")
    t1 = time.perf_counter()
    plain = code.memory_circuit.detector_error_model()
    try:
        dem = code.memory_circuit.detector_error_model(decompose_errors=True)
        decomposed = True
    except ValueError as exc:
        print("  decomposition failed -> undecomposed model (sinter's fallback):", str(exc)[:120])
        dem, decomposed = plain, False
    t2 = time.perf_counter()
    s = plain.summary()
    print(f"{name}: {code.name}  r={code.distance}  D={s['D']} M={s['M']} nnz(H)={s['nnz_H']} sum_p={s['sum_p']:.1f}  dets/mechanism {s['dets_per_mechanism']}")
    assert s["D"] == want_D and s["M"] == want_M_plain, (s["D"], s["M"])
    a, b, om, pp = dem.graphlike_components()
    ncomp = int(dem.comp_a.shape[0])
    dropped = int((dem.comp_a == 0xFFFFFFFF).sum())
    hyper = int((np.diff(plain.det_indptr.astype(np.int64)) > 2).sum())
    print(f"  decomposed={decomposed}: {dem.num_errors} errors, {ncomp} components, {dropped} not graphlike (dropped by the graph); "
          f"{hyper} of {plain.num_errors} symptoms flip more than two detectors; circuit {t1 - t0:.2f} s, DEMs {t2 - t1:.2f} s")
    demh = be.DemHandle(dem, device=0)
    t3 = time.perf_counter()
    dech = be.DecoderHandle(demh, be.MATCHING, with_corrections=True)
    t4 = time.perf_counter()
    gi = dech.graph_info
    print(f"  matching graph: {gi.num_edges} edges ({gi.num_boundary_edges} boundary), longest shortest path {gi.max_distance}; decoder set-up {t4 - t3:.1f} s")
    # graph == the oracle's own build from the DEM
    u, v, w, omk, pr = dech.edges()
    got = {(int(x), -1 if int(y) == be.NO_DET else int(y)): (float(q), int(o), int(ww)) for x, y, q, o, ww in zip(u, v, pr, omk, w)}
    mine = po.merged_edges(dem)
    iw = po.integer_weights(np.array([e[2] for e in mine]))
    ref = {(e[0], e[1]): (e[2], e[3], int(x)) for e, x in zip(mine, iw)}
    assert got.keys() == ref.keys()
    assert all(abs(got[k][0] - ref[k][0]) < 1e-12 and got[k][1:] == ref[k][1:] for k in got)
    print("  graph identical to the oracle's own build (edges, merged probabilities, observable masks, integer weights)")
    # a sample of shots against the exact-MWPM oracle built from the DEM
    orc = po.MwpmOracle(dem)
    dets, obs = demh.sample(7, 0, check_shots)
    pred, wts, flags, corr = dech.decode(dets, want_weights=True, want_flags=True, want_corrections=True)
    d8 = b8(dets, dem.num_detectors)
    t5 = time.perf_counter()
    opred, owts = orc.decode_batch(d8, nthreads=os.cpu_count() or 4, return_weights=True)
    t6 = time.perf_counter()
    assert (flags.cpu().numpy() == 0).all() and orc.last_failures == 0
    assert (wts.cpu().numpy() == owts).all(), "GPU matching weight differs from the oracle optimum"
    cb = np.unpackbits(corr.cpu().numpy().view(np.uint8).reshape(check_shots, -1), axis=1, bitorder="little")[:, : len(u)]
    syn = np.zeros((check_shots, dem.num_detectors), dtype=np.uint8)
    for e in range(len(u)):
        col = cb[:, e]
        syn[:, u[e]] ^= col
        if v[e] != be.NO_DET:
            syn[:, v[e]] ^= col
    want = np.unpackbits(d8, axis=1, bitorder="little")[:, : dem.num_detectors]
    assert (syn == want).all(), "a correction does not reproduce its syndrome"
    p32 = pred.cpu().numpy().view(np.uint32)[:, 0]
    lc = (cb.astype(np.int64) @ (omk & np.uint64(1)).astype(np.int64)) & 1
    assert (lc == (p32 & 1)).all()
    ndef = want.sum(axis=1)
    print(f"  {check_shots} shots ({ndef.mean():.0f} defects/shot, max {ndef.max()}): weights == oracle optimum, every correction reproduces its syndrome, "
          f"prediction == L.c; predictions equal to the oracle's on {(p32 == opred).mean() * 100:.1f} % (ties); oracle {check_shots / (t6 - t5):.0f} shots/s on {os.cpu_count()} threads")
    dech.close()
    dech = be.DecoderHandle(demh, be.MATCHING)
    dech.collect(1, 0, min(rate_shots, 1 << 14), 0)
    t7 = time.perf_counter()
    shots, errors, _ = dech.collect(11, 0, rate_shots, 0)
    t8 = time.perf_counter()
    lo, hi = gq.wilson_interval(errors, shots)
    print(f"  gq_collect: {shots} shots in {t8 - t7:.2f} s = {shots / (t8 - t7):.3e} shots/s; {errors} logical errors, LER {errors / shots:.3e} [{lo:.3e}, {hi:.3e}]")
    dech.close(); demh.close()


def steane(shots):
    H = np.array(H_STEANE)
    for p in (2e-4, 5e-4, 1e-3, 2e-3):
        code = gq.CssCode(Hx=H, Hz=H, name="Steane", depolarize1_rate=p, depolarize2_rate=p)
        code.build_memory_circuit(number_of_rounds=code.distance, logic_check="Z")
        try:
            dem = code.memory_circuit.detector_error_model(decompose_errors=True)
        except ValueError:
            dem = code.memory_circuit.detector_error_model()
        demh = be.DemHandle(dem, device=0)
        dech = be.DecoderHandle(demh, be.MATCHING)
        dech.collect(1, 0, 1 << 20, 0)
        t0 = time.perf_counter()
        n, e, _ = dech.collect(20261019, 0, int(shots), 0)
        dt = time.perf_counter() - t0
        lo, hi = gq.wilson_interval(e, n)
        print(f"steane p={p:g}: {n} shots in {dt:.2f} s = {n / dt:.3e} shots/s; {e} errors, LER {e / n:.4e} [{lo:.4e}, {hi:.4e}]  LER/p = {e / n / p:.2f}", flush=True)
        dech.close(); demh.close()


def bb144(shots):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
    for p in (1e-3, 2e-3, 3e-3):
        th = gq.ThresholdLAB(code=gq.BivariateBicycleCode, configurations=[dict(BB144)], error_rates=[p], decoder="bposd")
        t0 = time.perf_counter()
        th.collect_stats(max_shots=int(shots), max_errors=10**9, logic_check="Z")
        dt = time.perf_counter() - t0
        st = th.samples[0]
        if rank == 0:
            print(f"bb144 p={p:g} ranks={world}: {st.json_metadata['name']} r={st.json_metadata['r']}: {st.shots} shots, {st.errors} errors, {st.discards} discards, "
                  f"LER {st.errors / max(1, st.shots - st.discards):.3e}; {dt:.1f} s incl. set-up = {st.shots / dt:.3e} shots/s whole job", flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    what = sys.argv[1]
    if what == "toric15":
        p = float(sys.argv[2]) if len(sys.argv) > 2 else 0.005
        hx, hz = toric_checks(15, 15)
        code = gq.CssCode(Hx=hx, Hz=hz, name="Toric", depolarize1_rate=p, depolarize2_rate=p)
        matching_config("toric15", code, 6750, 147150, 128, 1 << 17)
    elif what == "unrot15":
        p = float(sys.argv[2]) if len(sys.argv) > 2 else 0.005
        code = gq.UnrotatedSurfaceCode(distance=15, depolarize1_rate=p, depolarize2_rate=p)
        matching_config("unrot15", code, 6300, 129831, 128, 1 << 17)
    elif what == "rot11":
        code = gq.RotatedSurfaceCode(distance=11, depolarize1_rate=0.005, depolarize2_rate=0.005)
        matching_config("rot11", code, 1320, 24483, 2048, 1 << 22)
    elif what == "steane":
        steane(float(sys.argv[2]) if len(sys.argv) > 2 else 1e9)
    elif what == "bb144":
        bb144(float(sys.argv[2]) if len(sys.argv) > 2 else 65536)
    else:
        raise SystemExit(__doc__)


if __name__ == "__main__":
    main()
