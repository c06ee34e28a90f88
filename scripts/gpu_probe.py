"""First-contact GPU probe: sampler/extractor parity, decoder optimality vs the oracle, timing."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from graphqec_b200 import RotatedSurfaceCode, RepetitionCode
from graphqec_b200 import backend as be
from oracle import pyoracle as po

def sm_to_b8(t, nbits):
    a = t.cpu().numpy().view(np.uint8).reshape(t.shape[0], -1)
    return np.ascontiguousarray(a[:, : (nbits + 7) // 8])

def run(d, p, nshots, big):
    code = RotatedSurfaceCode(distance=d, depolarize1_rate=p, depolarize2_rate=p)
    code.build_memory_circuit(number_of_rounds=d, logic_check="Z")
    dem = code.memory_circuit.detector_error_model(decompose_errors=True)
    print(f"d={d} p={p}", dem.summary()["M"], flush=True)
    h = be.DemHandle(dem)
    print(" tile", h.tile_shots, "classes", h.info.num_classes, "segments", h.info.num_segments, "sum_p", h.info.sum_p)
    t0 = time.time()
    dec = be.DecoderHandle(h, be.MATCHING, with_corrections=True)
    gi = dec.graph_info
    print(" decoder create %.2fs" % (time.time() - t0), "edges", gi.num_edges, "bnd", gi.num_boundary_edges, "fast_cap", gi.fast_capacity, "maxdist", gi.max_distance, flush=True)
    dets, obs, errs = h.sample(1234, 0, nshots, return_errors=True)
    # ---- extraction parity on identical error vectors
    dets_dm, obs_dm = h.extract(errs, nshots)
    dets_sm2 = h.dm_to_sm(dets_dm, h.D, nshots)
    obs_sm2 = h.dm_to_sm(obs_dm, max(h.O, 1), nshots)
    print(" extract==sample dets:", bool((dets_sm2 == dets).all()), "obs:", bool((obs_sm2 == obs).all()))
    back = h.sm_to_dm(dets, h.D, nshots)
    print(" transpose roundtrip:", bool((back == dets_dm).all()))
    # oracle H.e on the same errors
    e_bits = np.unpackbits(errs.cpu().numpy().view(np.uint8).reshape(errs.shape[0], -1), axis=1, bitorder="little")[: dem.num_errors, :nshots]
    od, oo = po.extract(dem, np.ascontiguousarray(e_bits.T))
    print(" oracle H.e == gpu:", bool((od == sm_to_b8(dets, h.D)).all()), bool((oo == sm_to_b8(obs, h.O)).all()))
    rate = e_bits.mean(axis=1)
    print(" mean fired/shot %.3f (sum_p %.3f)" % (e_bits.sum() / nshots, dem.probs.sum()), "max |rate-p|/sigma %.2f" % np.max(np.abs(rate - dem.probs) / np.sqrt(dem.probs * (1 - dem.probs) / nshots + 1e-30)))
    # ---- decode
    torch.cuda.synchronize(); t0 = time.time()
    pred, wts, flags, corr = dec.decode(dets, want_weights=True, want_flags=True, want_corrections=True)
    torch.cuda.synchronize(); dt = time.time() - t0
    print(" decode %d shots in %.3fs (%.0f shots/s) flags!=0: %d" % (nshots, dt, nshots / dt, int((flags != 0).sum())), flush=True)
    u, v, w, om, pr = dec.edges()
    edges = [(int(a), -1 if int(b) == be.NO_DET else int(b), float(q), int(o)) for a, b, q, o in zip(u, v, pr, om)]
    orc = po.MwpmOracle(edges=edges, num_detectors=dem.num_detectors, num_observables=dem.num_observables, int_weights=w.astype(np.int64))
    nchk = min(nshots, 2000)
    b8 = sm_to_b8(dets, h.D)
    opred, owts = orc.decode_batch(b8[:nchk], nthreads=8, return_weights=True)
    gw = wts.cpu().numpy()[:nchk]
    print(" weight parity vs oracle: %d/%d equal" % (int((gw == owts).sum()), nchk), "pred equal frac %.4f" % float((pred.cpu().numpy().view(np.uint32)[:nchk, 0] == opred).mean()))
    # independent graph build
    mine = po.merged_edges(dem)
    print(" edge list == oracle's own build:", [(e[0], e[1], e[3]) for e in mine] == [(e[0], e[1], e[3]) for e in edges],
          "max |dp| %.2e" % max(abs(a[2] - b[2]) for a, b in zip(mine, edges)),
          "weights equal:", bool((po.integer_weights(np.array([e[2] for e in mine])) == w).all()))
    # ---- syndrome reproduction of corrections
    cb = np.unpackbits(corr.cpu().numpy().view(np.uint8).reshape(nshots, -1), axis=1, bitorder="little")[:, : len(u)]
    H = np.zeros((len(u), dem.num_detectors), dtype=np.uint8)
    for k, (a, b) in enumerate(zip(u, v)):
        H[k, a] ^= 1
        if b != be.NO_DET: H[k, b] ^= 1
    syn = (cb[:nchk].astype(np.int32) @ H.astype(np.int32)) & 1
    db = np.unpackbits(b8[:nchk], axis=1, bitorder="little")[:, : dem.num_detectors]
    print(" corrections reproduce syndrome:", bool((syn == db).all()))
    lobs = ((cb[:nchk].astype(np.int64) @ (om & 1).astype(np.int64)) & 1)
    print(" L.c == prediction:", bool((lobs == (pred.cpu().numpy().view(np.uint32)[:nchk, 0] & 1)).all()))
    gerr = h.count_errors(pred, obs)
    print(" gpu errors %d / %d" % (gerr, nshots), " oracle errors on first %d: %d, gpu %d" % (nchk, po.count_errors(opred, sm_to_b8(obs, h.O)[:nchk]), int((pred.cpu().numpy().view(np.uint32)[:nchk, 0] != obs.cpu().numpy().view(np.uint32)[:nchk, 0]).sum())))
    if big:
        t0 = time.time()
        res = dec.collect(20261019, 0, big, 0)
        dt = time.time() - t0
        print(" collect", res, "%.3fs -> %.3e shots/s" % (dt, big / dt), dec.last_timing(), flush=True)

if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    run(3, 0.01, 4096, 0)
    run(5, 0.005, 8192, 1 << 20)
    run(11, 0.005, 8192, 1 << 20)
