"""GPU parity tests proper: every call goes through the C ABI (libgqec_b200.so) and is checked against
the CPU oracle on the same inputs.  North-star tolerances (BASELINE.json):
 (1) detectors/observables bit-exact for identical error vectors,
 (2) every correction reproduces its shot's syndrome exactly,
 (3) logical error rate inside the oracle's 95 % Wilson interval.
Additionally (stronger than the north star asks): the matching found is *optimal* shot by shot --
its integer weight equals the exact-MWPM oracle's.
"""
import json
import os

import numpy as np
import pytest

import graphqec_b200 as gq
from graphqec_b200 import backend as be
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu

STEANE_H = np.array([[1, 1, 1, 1, 0, 0, 0], [0, 1, 1, 0, 1, 1, 0], [0, 0, 1, 1, 0, 1, 1]])


def _toric(L):
    n = 2 * L * L
    Hx = np.zeros((L * L, n), dtype=int)
    Hz = np.zeros((L * L, n), dtype=int)
    h = lambda x, y: (x % L) * L + (y % L)
    v = lambda x, y: L * L + (x % L) * L + (y % L)
    for x in range(L):
        for y in range(L):
            Hx[x * L + y, [h(x, y), h(x, y - 1), v(x, y), v(x - 1, y)]] = 1
            Hz[x * L + y, [h(x, y), h(x + 1, y), v(x, y), v(x, y + 1)]] = 1
    return Hx, Hz


def make_dem(kind, p, **kw):
    if kind == "steane":
        code = gq.CssCode(Hx=STEANE_H, Hz=STEANE_H, depolarize1_rate=p, depolarize2_rate=p)
    elif kind == "toric":
        Hx, Hz = _toric(kw.pop("L"))
        code = gq.CssCode(Hx=Hx, Hz=Hz, name="Toric", depolarize1_rate=p, depolarize2_rate=p)
    else:
        lc = kw.pop("logic_check", "Z")
        code = getattr(gq, kind)(**kw, depolarize1_rate=p, depolarize2_rate=p)
        code.build_memory_circuit(number_of_rounds=kw.get("rounds", code.distance), logic_check=lc)
        try:
            return code.memory_circuit.detector_error_model(decompose_errors=True)
        except ValueError:
            return code.memory_circuit.detector_error_model()
    code.build_memory_circuit(number_of_rounds=kw.get("rounds", code.distance), logic_check="Z")
    try:
        return code.memory_circuit.detector_error_model(decompose_errors=True)
    except ValueError:
        return code.memory_circuit.detector_error_model()


CASES = {
    "rep3": ("RepetitionCode", 0.1, {"distance": 3}),
    "rep7": ("RepetitionCode", 0.1, {"distance": 7}),
    "rot3": ("RotatedSurfaceCode", 0.01, {"distance": 3}),
    "rot5": ("RotatedSurfaceCode", 0.005, {"distance": 5}),
    "rot7": ("RotatedSurfaceCode", 0.008, {"distance": 7}),
    "rot11": ("RotatedSurfaceCode", 0.005, {"distance": 11}),
    "unrot5": ("UnrotatedSurfaceCode", 0.005, {"distance": 5}),
    "rot4_even": ("RotatedSurfaceCode", 0.005, {"distance": 4}),
    "rot9_hi": ("RotatedSurfaceCode", 0.012, {"distance": 9}),        # ~110 defects: K=4 tier, K=8 stragglers
    "rot11_hi": ("RotatedSurfaceCode", 0.01, {"distance": 11}),       # ~157 defects: K=8 tier
    "unrot7": ("UnrotatedSurfaceCode", 0.008, {"distance": 7}),      # > 255 defects for some shots: memory tier
    "steane": ("steane", 0.003, {}),
    "toric4": ("toric", 0.004, {"L": 4}),
    "rot5_X": ("RotatedSurfaceCode", 0.005, {"distance": 5, "logic_check": "X"}),
    "unrot5_X": ("UnrotatedSurfaceCode", 0.006, {"distance": 5, "logic_check": "X"}),
    "rot6_even": ("RotatedSurfaceCode", 0.004, {"distance": 6}),
}
_cache = {}


def case(name):
    if name not in _cache:
        kind, p, kw = CASES[name]
        dem = make_dem(kind, p, **dict(kw))
        demh = be.DemHandle(dem, device=0)
        dech = be.DecoderHandle(demh, be.MATCHING, with_corrections=True)
        _cache[name] = (dem, demh, dech)
    return _cache[name]


def b8(t, nbits):
    a = t.cpu().numpy().view(np.uint8).reshape(t.shape[0], -1)
    return np.ascontiguousarray(a[:, : max(1, (nbits + 7) // 8)])


def bits_mm(errs, M, nshots):
    a = errs.cpu().numpy().view(np.uint8).reshape(errs.shape[0], -1)
    return np.unpackbits(a, axis=1, bitorder="little")[:M, :nshots]


def oracle_for(dem, dech):
    """The exact-MWPM oracle, built from the DEM by the oracle's own graph builder (merged edges, first observable
    mask, integer weights) -- NOT from the GPU decoder's edge list; the edge arrays are returned for the correction
    checks only.  test_matching_graph_equals_oracles_own_build compares the two graphs edge by edge."""
    return po.MwpmOracle(dem), dech.edges()


# ------------------------------------------------------------------------------------------ K1 / K2
@pytest.mark.parametrize("name,nshots", [("rep3", 1000), ("rot5", 4099), ("rot11", 2048), ("unrot5", 777), ("steane", 5000), ("toric4", 1500)])
def test_extraction_bit_exact_on_identical_error_vectors(name, nshots):
    """Tolerance (1): gq_sample's detectors/observables == gq_extract(errors) == oracle H.e, L.e."""
    dem, demh, _ = case(name)
    dets, obs, errs = demh.sample(99, 0, nshots, return_errors=True)
    dets_dm, obs_dm = demh.extract(errs, nshots)
    assert (demh.dm_to_sm(dets_dm, demh.D, nshots) == dets).all()
    assert (demh.dm_to_sm(obs_dm, max(demh.O, 1), nshots) == obs).all()
    e = bits_mm(errs, dem.num_errors, nshots)
    od, oo = po.extract(dem, np.ascontiguousarray(e.T))
    assert (od == b8(dets, dem.num_detectors)).all()
    assert (oo == b8(obs, dem.num_observables)).all()


def test_extract_random_dense_errors_and_transpose_roundtrip():
    import torch
    dem, demh, _ = case("rot7")
    nshots = 3001
    rng = np.random.default_rng(4)
    e = (rng.random((dem.num_errors, nshots)) < 0.3).astype(np.uint8)
    sw = (nshots + 31) // 32
    packed = np.zeros((dem.num_errors, sw * 32), dtype=np.uint8)
    packed[:, :nshots] = e
    words = np.packbits(packed, axis=1, bitorder="little").view(np.uint32)
    errs = torch.from_numpy(words.view(np.int32).copy()).cuda()
    dets_dm, obs_dm = demh.extract(errs, nshots)
    od, oo = po.extract(dem, np.ascontiguousarray(e.T))
    sm = demh.dm_to_sm(dets_dm, demh.D, nshots)
    assert (b8(sm, dem.num_detectors) == od).all()
    assert (b8(demh.dm_to_sm(obs_dm, 1, nshots), 1) == oo).all()
    assert (demh.sm_to_dm(sm, demh.D, nshots) == dets_dm).all()


def test_sampler_mechanism_rates():
    """Every mechanism fires with its own probability (z-score per probability class)."""
    dem, demh, _ = case("rot5")
    nshots = 1 << 18
    _, _, errs = demh.sample(5, 0, nshots, return_errors=True)
    fired = bits_mm(errs, dem.num_errors, nshots).sum(axis=1)
    keys = np.round(dem.probs, 12)
    for p in np.unique(keys):
        sel = keys == p
        n = sel.sum() * nshots
        z = (fired[sel].sum() - n * p) / np.sqrt(n * p * (1 - p))
        assert abs(z) < 4.5, (p, z)
    # and no mechanism is starved or duplicated
    z1 = (fired - nshots * dem.probs) / np.sqrt(nshots * dem.probs * (1 - dem.probs))
    assert np.abs(z1).max() < 6


def test_sampler_is_counter_based():
    """Same (seed, shot range) -> same bits, whatever the batching: basis of GPU-count independence."""
    dem, demh, _ = case("rot5")
    t = demh.tile_shots
    a, ao = demh.sample(42, 0, 8 * t)
    b1, bo1 = demh.sample(42, 0, 3 * t)
    b2, bo2 = demh.sample(42, 3 * t, 5 * t)
    assert (a[: 3 * t] == b1).all() and (a[3 * t:] == b2).all()
    assert (ao[: 3 * t] == bo1).all() and (ao[3 * t:] == bo2).all()
    c, _ = demh.sample(43, 0, 8 * t)
    assert not (a == c).all()
    # ragged tail: a partial tile equals the head of the full one
    d, _ = demh.sample(42, 0, 3 * t + 17)
    assert (d == a[: 3 * t + 17]).all()
    with pytest.raises(be.GqError):
        demh.sample(42, 5, 10)


def test_detector_rates_match_dem_marginals():
    dem, demh, _ = case("rot11")
    nshots = 1 << 17
    dets, _ = demh.sample(3, 0, nshots)
    got = np.unpackbits(b8(dets, dem.num_detectors), axis=1, bitorder="little")[:, : dem.num_detectors].mean(axis=0)
    keep = np.ones(dem.num_detectors)
    np.multiply.at(keep, dem.det_idx, np.repeat(1 - 2 * dem.probs, np.diff(dem.det_indptr.astype(np.int64))))
    want = 0.5 * (1 - keep)
    assert np.abs(got - want).max() < 6 * np.sqrt(want.max() / nshots)
    assert abs(got.sum() - 84.06) < 0.5          # 84.06 detection events per shot


# ------------------------------------------------------------------------------------------ K3a
@pytest.mark.parametrize("name,nshots", [("rep3", 3000), ("rep7", 3000), ("rot3", 3000), ("rot5", 3000), ("rot7", 2000),
                                         ("rot11", 1500), ("unrot5", 1500), ("rot4_even", 2000), ("steane", 3000), ("toric4", 1500),
                                         ("rot9_hi", 1000), ("rot11_hi", 800), ("unrot7", 600), ("rot5_X", 1500), ("unrot5_X", 1000),
                                         ("rot6_even", 1500)])
def test_matching_is_optimal_and_corrections_reproduce_syndrome(name, nshots):
    dem, demh, dech = case(name)
    orc, (u, v, w, om, pr) = oracle_for(dem, dech)
    dets, obs = demh.sample(7, 0, nshots)
    pred, wts, flags, corr = dech.decode(dets, want_weights=True, want_flags=True, want_corrections=True)
    d8 = b8(dets, dem.num_detectors)
    opred, owts = orc.decode_batch(d8, nthreads=8, return_weights=True)
    flags = flags.cpu().numpy()
    ok = flags == 0
    # shots the oracle can match, the GPU must match too -- with the same (minimum) weight
    matched = owts < (1 << 29)
    if name not in ("steane", "toric4"):
        assert ok.all() and orc.last_failures == 0
    assert (wts.cpu().numpy()[ok] == owts[ok]).all()
    # tolerance (2): H_graph . c == syndrome, and the prediction is L . c
    cb = np.unpackbits(corr.cpu().numpy().view(np.uint8).reshape(nshots, -1), axis=1, bitorder="little")[:, : len(u)]
    H = np.zeros((len(u), dem.num_detectors), dtype=np.int32)
    H[np.arange(len(u)), u] ^= 1
    inner = v != be.NO_DET
    H[np.flatnonzero(inner), v[inner]] ^= 1
    syn = (cb.astype(np.int32) @ H) & 1
    want = np.unpackbits(d8, axis=1, bitorder="little")[:, : dem.num_detectors]
    assert (syn[ok] == want[ok]).all()
    p32 = pred.cpu().numpy().view(np.uint32)[:, 0]
    for bit in range(dem.num_observables):
        lc = (cb.astype(np.int64) @ ((om >> np.uint64(bit)) & np.uint64(1)).astype(np.int64)) & 1
        assert (lc[ok] == ((p32[ok] >> bit) & 1)).all()
    # predictions agree with the oracle except where minimum-weight matchings tie
    assert (p32[ok] == opred[ok]).mean() > 0.97


def test_matching_graph_equals_oracles_own_build():
    """A8 conventions: merged edges, probabilities and integer weights, built independently -- every case of this file:
    repetition, rotated (odd and even d, Z and X memory), unrotated (Z and X), toric and Steane through CssCode."""
    for name in CASES:
        dem, demh, dech = case(name)
        u, v, w, om, pr = dech.edges()
        got = {(int(a), -1 if int(b) == be.NO_DET else int(b)): (float(q), int(o), int(ww)) for a, b, q, o, ww in zip(u, v, pr, om, w)}
        mine = po.merged_edges(dem)
        iw = po.integer_weights(np.array([e[2] for e in mine]))
        ref = {(e[0], e[1]): (e[2], e[3], int(x)) for e, x in zip(mine, iw)}
        assert got.keys() == ref.keys()
        for k in got:
            assert abs(got[k][0] - ref[k][0]) < 1e-12 and got[k][1:] == ref[k][1:]


def test_decode_edge_cases():
    import torch
    dem, demh, dech = case("rot5")
    # empty batch, all-zero syndromes, single shot
    z = torch.zeros((0, demh.DW), dtype=torch.int32, device="cuda")
    assert dech.decode(z).shape[0] == 0
    z = torch.zeros((5, demh.DW), dtype=torch.int32, device="cuda")
    pred, wts = dech.decode(z, want_weights=True)
    assert (pred == 0).all() and (wts == 0).all()
    # every detector fired: far beyond the fast tier -> overflow tier must give the optimum too
    full = np.zeros((3, demh.DW), dtype=np.uint32)
    for k in range(dem.num_detectors):
        full[:, k >> 5] |= np.uint32(1) << np.uint32(k & 31)
    full[1, 0] ^= 1   # odd defect count
    t = torch.from_numpy(full.view(np.int32)).cuda()
    pred, wts, flags = dech.decode(t, want_weights=True, want_flags=True)
    orc, _ = oracle_for(dem, dech)
    opred, owts = orc.decode_batch(b8(t, dem.num_detectors), return_weights=True)
    assert (flags == 0).all() and (wts.cpu().numpy() == owts).all()


def test_decode_b8_host_path_equals_device_path():
    """gq_decode_b8 == sinter's decode_shots_bit_packed layout: host uint8 rows in, uint8 rows out."""
    dem, demh, dech = case("rot7")
    dets, obs = demh.sample(21, 0, 5000)
    pred = dech.decode(dets)
    out = dech.decode_b8(b8(dets, dem.num_detectors))
    assert out.shape == (5000, 1) and out.dtype == np.uint8
    assert (out == b8(pred, dem.num_observables)).all()
    with pytest.raises(ValueError):
        dech.decode_b8(np.zeros((4, 3), dtype=np.uint8))


def test_decode_b8_pipelined_chunks_equal_device_path():
    """a host batch long enough to be cut into several chunks (H2D of chunk c+1 overlaps the decode of c)
    decodes exactly like the device-resident path, ragged tail included"""
    dem, demh, dech = case("rep7")
    n = (1 << 19) + 64 * demh.tile_shots + 0   # > 2^19 shots: gq_decode_b8 cuts it into 2^18-shot chunks
    dets, obs = demh.sample(33, 0, n)
    n_ragged = n - 77
    pred = dech.decode(dets[:n_ragged])
    out = dech.decode_b8(b8(dets[:n_ragged], dem.num_detectors))
    assert out.shape == (n_ragged, 1)
    assert (out == b8(pred, dem.num_observables)).all()


def test_decode_b8_growing_chunk_schedule_equals_device_path():
    """>= 2^21 shots: chunks of 2^18, 2^19, 2^20, 2^21 ... shots, a short remainder riding with the last one"""
    dem, demh, dech = case("rep7")
    n = (1 << 22) + (1 << 18) + 5 * demh.tile_shots
    dets, obs = demh.sample(34, 0, n)
    n_ragged = n - 13
    pred = dech.decode(dets[:n_ragged])
    out = dech.decode_b8(b8(dets[:n_ragged], dem.num_detectors))
    assert out.shape == (n_ragged, 1)
    assert (out == b8(pred, dem.num_observables)).all()


def test_every_first_tier_class_and_the_ladder_give_the_optimum():
    """syndromes with 1 .. 1200 defects: classes K = 1..8 (8-bit vertex keys), K = 10, 12, 16 (9-bit keys) and
    K = 24, 32 (10-bit keys) of the first tier, then the older tier"""
    import torch
    dem, demh, dech = case("rot11")
    rng = np.random.default_rng(3)
    D = dem.num_detectors
    counts = [1, 2, 31, 32, 33, 64, 65, 96, 97, 128, 129, 160, 161, 192, 193, 224, 225, 255, 256, 257, 300, 320, 321,
              384, 385, 450, 512, 513, 700, 768, 769, 1000, 1024, 1025, 1200]
    rows = np.zeros((len(counts) * 2, demh.DW), dtype=np.uint32)
    for i, c in enumerate(counts * 2):
        for k in rng.choice(D, size=min(c, D), replace=False):
            rows[i, k >> 5] |= np.uint32(1) << np.uint32(k & 31)
    t = torch.from_numpy(rows.view(np.int32)).cuda()
    pred, wts, flags = dech.decode(t, want_weights=True, want_flags=True)
    orc, _ = oracle_for(dem, dech)
    opred, owts = orc.decode_batch(b8(t, D), return_weights=True)
    assert (flags == 0).all() and (wts.cpu().numpy() == owts).all()
    # the throughput instantiation (no weights / flags requested) must predict the same
    assert (dech.decode(t) == pred).all()


def test_first_tier_classes_up_to_2048_defects():
    """classes K = 48, 64 (11-bit vertex keys, 1025 .. 2048 defects per shot: rotated d = 23 up to p = 1e-2, the largest
    point of the reference's example) and the ladder beyond them, on random syndromes of the rotated d = 17 graph"""
    import torch
    code = gq.RotatedSurfaceCode(distance=17, depolarize1_rate=0.005, depolarize2_rate=0.005)
    code.build_memory_circuit(number_of_rounds=17, logic_check="Z")
    dem = code.memory_circuit.detector_error_model(decompose_errors=True)
    demh = be.DemHandle(dem, device=0)
    dech = be.DecoderHandle(demh, be.MATCHING)
    assert dech.graph_info.max_distance < 32768
    rng = np.random.default_rng(5)
    D = dem.num_detectors
    counts = [1024, 1025, 1300, 1536, 1537, 1800, 2047, 2048, 2049, 2300]
    rows = np.zeros((len(counts), demh.DW), dtype=np.uint32)
    for i, c in enumerate(counts):
        for k in rng.choice(D, size=c, replace=False):
            rows[i, k >> 5] |= np.uint32(1) << np.uint32(k & 31)
    t = torch.from_numpy(rows.view(np.int32)).cuda()
    pred, wts, flags = dech.decode(t, want_weights=True, want_flags=True)
    orc = po.MwpmOracle(dem)
    opred, owts = orc.decode_batch(b8(t, D), nthreads=len(counts), return_weights=True)
    assert (flags == 0).all() and (wts.cpu().numpy() == owts).all()
    assert (dech.decode(t) == pred).all()
    dech.close(); demh.close()


def test_count_and_collect_agree():
    dem, demh, dech = case("rot5")
    n = 40 * demh.tile_shots
    dets, obs = demh.sample(1234, 0, n)
    pred = dech.decode(dets)
    want = int((pred != obs).any(dim=1).sum().item())
    assert demh.count_errors(pred, obs) == want
    assert dech.collect(1234, 0, n, 0) == (n, want, 0)
    # split across "ranks": the same shots in two calls give the same total
    a = dech.collect(1234, 0, 16 * demh.tile_shots, 0)
    b = dech.collect(1234, 16 * demh.tile_shots, 24 * demh.tile_shots, 0)
    assert a[1] + b[1] == want


def test_collect_early_stop_on_max_errors():
    dem, demh, dech = case("rep3")
    shots, errors, _ = dech.collect(1, 0, 1 << 23, 1000)
    assert errors >= 1000 and shots < (1 << 23) and shots % (1 << 20) == 0


# ------------------------------------------------------------------------------------------ LER
def test_logical_error_rates_inside_oracle_wilson_intervals(golden_dir):
    """Tolerance (3), at matched shot counts, against fixtures made by tests/golden/make_oracle_ler.py."""
    fixtures = json.load(open(os.path.join(golden_dir, "oracle_ler.json")))
    assert len(fixtures) >= 10
    for fx in fixtures:
        kwargs = {k: (np.array(v) if isinstance(v, list) else v) for k, v in fx["kwargs"].items()}
        code = getattr(gq, fx["code"])(**kwargs, depolarize1_rate=fx["p"], depolarize2_rate=fx["p"])
        code.build_memory_circuit(number_of_rounds=fx["rounds"], logic_check=fx["logic_check"])
        try:
            dem = code.memory_circuit.detector_error_model(decompose_errors=True)
        except ValueError:           # sinter's rule: fall back to the undecomposed model (Steane through CssCode)
            dem = code.memory_circuit.detector_error_model()
        demh = be.DemHandle(dem, device=0)
        dech = be.DecoderHandle(demh, be.MATCHING)
        # the GPU draws many more shots than the fixture holds (16 x, at least 2^22), so that its own sampling noise is a
        # quarter of the oracle's: its estimate must lie inside the oracle's 95 % interval widened by that remainder
        n = max(1 << 22, 16 * fx["shots"])
        n = -(-n // demh.tile_shots) * demh.tile_shots
        shots, errors, _ = dech.collect(20261019, 0, n, 0)
        lo, hi = fx["wilson95"]
        glo, ghi = po.wilson(errors, shots)
        width = (ghi - glo) / 2
        assert lo - width <= errors / shots <= hi + width, (fx["code"], fx["kwargs"], fx["p"], errors / shots, lo, hi)
        dech.close(); demh.close()


def test_north_star_dem_on_device():
    dem, demh, dech = case("rot11")
    assert (demh.D, demh.O, demh.M) == (1320, 1, 27680)      # decomposed model: one error per distinct decomposition
    assert abs(demh.info.sum_p - 36.94) < 0.01 and 36 <= demh.info.num_classes <= 44
    gi = dech.graph_info
    assert gi.num_edges == 6930 and gi.num_boundary_edges == 360


# ------------------------------------------------------------------------------------------ K3b
def _bp_case(name, **params):
    dem, demh, _ = case(name)
    return dem, demh, be.DecoderHandle(demh, be.BP_MINSUM, **params)


@pytest.mark.parametrize("name,nshots,scratch", [("rot3", 300, False), ("steane", 300, False), ("rot5", 120, False),
                                                 ("rot3", 301, True), ("rot5", 123, True), ("toric4", 77, True),
                                                 ("rot3", 301, 2), ("steane", 300, 2), ("rot5", 123, 2), ("toric4", 77, 2)])
def test_bp_matches_oracle_bit_for_bit(name, nshots, scratch):
    """Same schedule, same fp32 summation order as oracle/bp_oracle.c: identical decisions and iteration counts -- for the
    one-shot-per-block kernel (messages in shared memory), for the shot-interleaved kernel (bp_scratch = 1 forces it here;
    ragged shot counts leave its last group of eight partly empty) and for the compressed min-sum kernel that big models
    get by default (bp_scratch = 2: check records + one sign bit per edge instead of messages)."""
    dem, demh, bp = _bp_case(name, bp_max_iter=25, bp_alpha=0.9, bp_scratch=scratch)
    orc = po.BpOracle(dem, max_iter=25, alpha=0.9)
    dets, obs = demh.sample(17, 0, nshots)
    pred, iters, flags, corr = bp.decode(dets, want_weights=True, want_flags=True, want_corrections=True)
    bits = np.unpackbits(b8(dets, dem.num_detectors), axis=1, bitorder="little")[:, : dem.num_detectors]
    cb = np.unpackbits(corr.cpu().numpy().view(np.uint8).reshape(nshots, -1), axis=1, bitorder="little")[:, : dem.num_errors]
    H, L = dem.dense_check_matrices()
    it_gpu = iters.cpu().numpy()
    nconv = 0
    for s in range(nshots):
        e, it = orc.decode(bits[s])
        assert it == it_gpu[s], (s, it, it_gpu[s])
        assert (e == cb[s]).all()
        if it > 0:
            nconv += 1
            assert ((H @ cb[s]) % 2 == bits[s]).all()          # converged output reproduces the syndrome
    p32 = pred.cpu().numpy().view(np.uint32)[:, 0]
    assert ((cb @ L.T) % 2 == (p32[:, None] & 1)).all()          # prediction = L.e
    assert ((flags.cpu().numpy() == 0) == (it_gpu > 0)).all()
    assert nconv > 0.5 * nshots


@pytest.mark.parametrize("name,nshots,iters,scratch", [("rot3", 400, 4, False), ("steane", 400, 3, False), ("rot5", 150, 6, False),
                                                       ("toric4", 150, 8, False), ("steane", 403, 3, True), ("rot5", 149, 6, True),
                                                       ("steane", 403, 3, 2), ("rot5", 149, 6, 2), ("toric4", 150, 8, 2)])
def test_bp_osd_matches_oracle_bit_for_bit(name, nshots, iters, scratch):
    """BP (few iterations, so that many shots stay unconverged) + OSD-0: the same error vector as
    oracle/bp_oracle.c::orc_bp_osd0 on every shot, and every solved shot reproduces its syndrome."""
    dem, demh, _ = case(name)
    bp = be.DecoderHandle(demh, be.BP_MINSUM, bp_max_iter=iters, bp_alpha=0.9, bp_osd=True, bp_scratch=scratch)
    orc = po.BpOracle(dem, max_iter=iters, alpha=0.9)
    dets, obs = demh.sample(23, 0, nshots)
    pred, it_gpu, flags, corr = bp.decode(dets, want_weights=True, want_flags=True, want_corrections=True)
    bits = np.unpackbits(b8(dets, dem.num_detectors), axis=1, bitorder="little")[:, : dem.num_detectors]
    cb = np.unpackbits(corr.cpu().numpy().view(np.uint8).reshape(nshots, -1), axis=1, bitorder="little")[:, : dem.num_errors]
    H, L = dem.dense_check_matrices()
    it_gpu, flags = it_gpu.cpu().numpy(), flags.cpu().numpy()
    n_osd = 0
    for s in range(nshots):
        e, it, st, nb, used = orc.decode_osd(bits[s], max_candidates=-2048)      # the kernel's schedule: 2048 ordered, then all
        assert it == it_gpu[s]
        assert (e == cb[s]).all(), (s, it, st)
        assert flags[s] == (1 if st == 1 else 0)
        if st != 1:
            assert ((H @ cb[s]) % 2 == bits[s]).all()
        n_osd += st == 0
    assert n_osd >= 10
    p32 = pred.cpu().numpy().view(np.uint32)[:, 0]
    assert ((cb @ L.T) % 2 == (p32[:, None] & 1)).all()          # prediction = L.e, OSD shots included
    # the throughput path (no flags / corrections requested) predicts the same
    assert (bp.decode(dets) == pred).all()
    bp.close()


def test_bposd_decodes_the_bivariate_bicycle_code():
    """decoder="bposd" end to end on BB [[72,12,6]]: with OSD every shot gets a valid correction, and the logical
    error rate drops well below BP's alone"""
    code = gq.BivariateBicycleCode(L1=6, L2=6, a1=3, a2=1, a3=2, b1=3, b2=1, b3=2, depolarize1_rate=0.003, depolarize2_rate=0.003)
    code.build_memory_circuit(number_of_rounds=3, logic_check="Z")
    dem = code.memory_circuit.detector_error_model()
    demh = be.DemHandle(dem, device=0)
    n = 4 * demh.tile_shots
    dets, obs = demh.sample(9, 0, n)
    bp = be.DecoderHandle(demh, be.BP_MINSUM, bp_max_iter=30)
    osd = be.DecoderHandle(demh, be.BP_MINSUM, bp_max_iter=30, bp_osd=True)
    p0, _, f0 = bp.decode(dets, want_weights=True, want_flags=True)
    p1, _, f1, corr = osd.decode(dets, want_weights=True, want_flags=True, want_corrections=True)
    assert (f0 != 0).float().mean().item() > 0.02 and (f1 != 0).sum().item() == 0
    H, L = dem.dense_check_matrices()
    bits = np.unpackbits(b8(dets, dem.num_detectors), axis=1, bitorder="little")[:, : dem.num_detectors]
    cb = np.unpackbits(corr.cpu().numpy().view(np.uint8).reshape(n, -1), axis=1, bitorder="little")[:, : dem.num_errors]
    assert (((cb.astype(np.int32) @ H.T.astype(np.int32)) % 2) == bits).all()      # every correction reproduces its syndrome
    e0 = (p0 != obs).any(dim=1).float().mean().item()
    e1 = (p1 != obs).any(dim=1).float().mean().item()
    assert e1 < 0.5 * e0
    bp.close(); osd.close(); demh.close()


def test_bp_on_bivariate_bicycle_code():
    """BASELINE configs[4] in miniature: BB [[72,12,6]], hyperedge DEM, min-sum BP through the fused path."""
    code = gq.BivariateBicycleCode(L1=6, L2=6, a1=3, a2=1, a3=2, b1=3, b2=1, b3=2, depolarize1_rate=0.001, depolarize2_rate=0.001)
    code.build_memory_circuit(number_of_rounds=2, logic_check="Z")
    dem = code.memory_circuit.detector_error_model()
    assert np.diff(dem.det_indptr.astype(int)).max() > 2          # genuinely non-matchable
    demh = be.DemHandle(dem, device=0)
    bp = be.DecoderHandle(demh, be.BP_MINSUM, bp_max_iter=30)
    n = 8 * demh.tile_shots
    dets, obs = demh.sample(3, 0, n)
    pred, iters, flags = bp.decode(dets, want_weights=True, want_flags=True)
    conv = (flags == 0)
    assert conv.float().mean().item() > 0.6
    # converged shots at this low error rate are almost always decoded correctly
    good = ((pred == obs).all(dim=1) & conv).sum().item() / max(1, conv.sum().item())
    assert good > 0.95
    shots, errors, _ = bp.collect(3, 0, n, 0)
    assert shots == n and errors == int((pred != obs).any(dim=1).sum().item())
    bp.close(); demh.close()


# ------------------------------------------------------------------------------------------ edges of the input space
def test_empty_single_zero_and_saturated_batches():
    """0 shots, 1 shot, all-zero and all-one syndromes (padding bits beyond D set) through both decoders"""
    import torch
    dem, demh, dech = case("rot3")
    bp = be.DecoderHandle(demh, be.BP_MINSUM, bp_max_iter=2, bp_osd=True)
    empty = torch.zeros((0, demh.DW), dtype=torch.int32, device="cuda")
    assert tuple(dech.decode(empty).shape) == (0, demh.OW) and tuple(bp.decode(empty).shape) == (0, demh.OW)
    assert dech.decode_b8(np.zeros((0, (dem.num_detectors + 7) // 8), np.uint8)).shape == (0, 1)
    one, _ = demh.sample(1, 0, 1)
    assert tuple(dech.decode(one).shape) == (1, demh.OW)
    assert dech.collect(1, 0, 1, 0)[0] == 1 and bp.collect(1, 0, demh.tile_shots, 0)[0] == demh.tile_shots
    zeros = torch.zeros((5, demh.DW), dtype=torch.int32, device="cuda")
    assert not dech.decode(zeros).any() and not bp.decode(zeros).any()
    ones = torch.full((3, demh.DW), -1, dtype=torch.int32, device="cuda")
    pred, wts, flags = dech.decode(ones, want_weights=True, want_flags=True)
    orc, _ = oracle_for(dem, dech)
    _, owts = orc.decode_batch(b8(ones, dem.num_detectors), return_weights=True)
    assert (flags == 0).all() and (wts.cpu().numpy() == owts).all()
    pb, ib, fb, cb = bp.decode(ones, want_weights=True, want_flags=True, want_corrections=True)
    H, _ = dem.dense_check_matrices()
    bits = np.unpackbits(cb.cpu().numpy().view(np.uint8).reshape(3, -1), axis=1, bitorder="little")[:, : dem.num_errors]
    assert (fb == 0).all() and (((bits @ H.T) % 2) == 1).all()       # OSD found a correction for the all-ones syndrome
    bp.close()


# ------------------------------------------------------------------------------------------ full-size properties
def test_north_star_corrections_at_scale():
    """Size-independent properties at the BASELINE size (2^20 shots of the north-star DEM, far beyond what the
    oracle can decode in a test): every correction reproduces its shot's syndrome (H_graph . c == s), the prediction
    is L . c, the reported weight is the weight of the correction, and the throughput path predicts the same."""
    import torch
    dem, demh, _ = case("rot11")
    dech = be.DecoderHandle(demh, be.MATCHING, with_corrections=True)
    u, v, w, om, pr = dech.edges()
    E, D = len(u), dem.num_detectors
    dev = torch.device("cuda")
    inc = torch.zeros((E, D + 2), dtype=torch.float16, device=dev)         # columns D, D+1: observable bit, edge weight slot
    inc[torch.arange(E), torch.from_numpy(u.astype(np.int64))] = 1
    inner = np.flatnonzero(v != be.NO_DET)
    inc[torch.from_numpy(inner), torch.from_numpy(v[inner].astype(np.int64))] = 1
    inc[:, D] = torch.from_numpy((om & np.uint64(1)).astype(np.float16)).to(dev)
    wt_vec = torch.from_numpy(w.astype(np.float32)).to(dev)
    n, chunk = 1 << 20, 1 << 15
    bad_syn = bad_obs = bad_wt = bad_fast = flagged = 0
    shifts = torch.arange(32, device=dev, dtype=torch.int32)
    for s0 in range(0, n, chunk):
        dets, obs = demh.sample(77, s0, chunk)
        pred, wts, flags, corr = dech.decode(dets, want_weights=True, want_flags=True, want_corrections=True)
        flagged += int((flags != 0).sum().item())
        cbits = ((corr.unsqueeze(-1) >> shifts) & 1).reshape(chunk, -1)[:, :E]
        res = (cbits.to(torch.float16) @ inc).to(torch.int32)              # sums stay far below 2048: exact in fp16
        sbits = ((dets.unsqueeze(-1) >> shifts) & 1).reshape(chunk, -1)[:, :D]
        bad_syn += int(((res[:, :D] & 1) != sbits).any(dim=1).sum().item())
        bad_obs += int(((res[:, D] & 1) != (pred[:, 0] & 1)).sum().item())
        bad_wt += int(((cbits.to(torch.float32) @ wt_vec).to(torch.int64) != wts.to(torch.int64)).sum().item())
        bad_fast += int((dech.decode(dets) != pred).any(dim=1).sum().item())
    assert flagged == 0 and bad_syn == 0 and bad_obs == 0 and bad_wt == 0 and bad_fast == 0
    dech.close()


def test_osd_walks_past_the_ordered_candidates_when_it_must():
    """2100 likely decoys that all flip D0 fill the 2048 reliability-ordered candidate slots; the syndrome (D1) needs one
    of ten symmetric, unlikely mechanisms behind them (BP cannot break their symmetry): the kernel goes on through all
    variables in index order -- like the oracle's max_candidates = -2048 schedule, bit for bit -- instead of giving up."""
    import torch
    lines = ["error(0.45) D0"] * 2100 + ["error(0.01) D1"] * 10 + ["error(0.01) D1 D2 L0", "error(0.02) D2"]
    dem = gq.DetectorErrorModel.from_text("\n".join(lines))
    demh = be.DemHandle(dem, device=0)
    bp = be.DecoderHandle(demh, be.BP_MINSUM, bp_max_iter=3, bp_alpha=0.9, bp_osd=True)
    orc = po.BpOracle(dem, max_iter=3, alpha=0.9)
    rows = np.array([[0b010], [0b110], [0b011], [0b000]], dtype=np.uint32)
    t = torch.from_numpy(rows.view(np.int32)).cuda()
    pred, it_gpu, flags, corr = bp.decode(t, want_weights=True, want_flags=True, want_corrections=True)
    cb = np.unpackbits(corr.cpu().numpy().view(np.uint8).reshape(len(rows), -1), axis=1, bitorder="little")[:, : dem.num_errors]
    walked_past = 0
    for s, row in enumerate(rows[:, 0]):
        syn = np.array([(row >> k) & 1 for k in range(3)], dtype=np.uint8)
        e, it, st, nb, used = orc.decode_osd(syn, max_candidates=-2048)
        assert it == int(it_gpu[s]) and (e == cb[s]).all() and int(flags[s]) == (1 if st == 1 else 0)
        assert st != 1
        walked_past += used > 2048
    assert walked_past >= 2
    assert orc.decode_osd(np.array([0, 1, 0], dtype=np.uint8), max_candidates=2048)[2] == 1     # the capped walk gives up
    bp.close(); demh.close()


def test_bb144_bposd_corrections_at_scale():
    """BASELINE configs[4] at its full size (BB [[144,12,12]], 12 rounds: 1728 detectors, 67 752 mechanisms): BP + OSD-0
    gives every shot a correction that reproduces its syndrome (H . e == s), the prediction is L . e, and almost no
    shot fails logically at p = 1e-3 -- where BP alone leaves four shots in five without a valid correction."""
    import torch
    code = gq.BivariateBicycleCode(depolarize1_rate=0.001, depolarize2_rate=0.001)
    code.build_memory_circuit(number_of_rounds=12, logic_check="Z")
    dem = code.memory_circuit.detector_error_model()
    assert (dem.num_detectors, dem.num_errors) == (1728, 67752)
    demh = be.DemHandle(dem, device=0)
    osd = be.DecoderHandle(demh, be.BP_MINSUM, bp_max_iter=30, bp_osd=True)
    bp = be.DecoderHandle(demh, be.BP_MINSUM, bp_max_iter=30)
    n = 2048
    dets, obs = demh.sample(5, 0, n)
    _, _, f0 = bp.decode(dets, want_weights=True, want_flags=True)
    pred, iters, flags, corr = osd.decode(dets, want_weights=True, want_flags=True, want_corrections=True)
    assert (f0 != 0).float().mean().item() > 0.5 and int((flags != 0).sum().item()) == 0
    dev = torch.device("cuda")
    M, D = dem.num_errors, dem.num_detectors
    H = torch.zeros((M, D + 1), dtype=torch.float16, device=dev)
    rows = np.repeat(np.arange(M), np.diff(dem.det_indptr.astype(np.int64)))
    H[torch.from_numpy(rows), torch.from_numpy(dem.det_idx.astype(np.int64))] = 1
    lrows = np.repeat(np.arange(M), np.diff(dem.obs_indptr.astype(np.int64)))
    H[torch.from_numpy(lrows), D] = 1                                   # every logical goes to observable 0
    shifts = torch.arange(32, device=dev, dtype=torch.int32)
    ebits = ((corr.unsqueeze(-1) >> shifts) & 1).reshape(n, -1)[:, :M]
    res = (ebits.to(torch.float16) @ H).to(torch.int32)                 # at most a few hundred ones per row: exact
    sbits = ((dets.unsqueeze(-1) >> shifts) & 1).reshape(n, -1)[:, :D]
    assert int(((res[:, :D] & 1) != sbits).sum().item()) == 0
    assert int(((res[:, D] & 1) != (pred[:, 0] & 1)).sum().item()) == 0
    assert (pred != obs).any(dim=1).float().mean().item() < 0.01
    bp.close(); osd.close(); demh.close()


def test_bb144_at_p3e3_every_shot_gets_a_valid_correction():
    """Round 1 capped the OSD basis at 1024 columns and the walk at 2048 candidates: 0.5 % of the shots at p = 3e-3 kept
    BP's unconverged decision.  With the basis sized by D and the walk continuing past the ordered candidates, all do."""
    code = gq.BivariateBicycleCode(depolarize1_rate=0.003, depolarize2_rate=0.003)
    code.build_memory_circuit(number_of_rounds=12, logic_check="Z")
    dem = code.memory_circuit.detector_error_model()
    demh = be.DemHandle(dem, device=0)
    osd = be.DecoderHandle(demh, be.BP_MINSUM, bp_max_iter=30, bp_osd=True)
    n = 4096
    dets, obs = demh.sample(9, 0, n)
    pred, iters, flags = osd.decode(dets, want_weights=True, want_flags=True)
    assert int((flags != 0).sum().item()) == 0
    shots, errors, discards = osd.collect(9, 0, n, 0)
    assert (shots, discards) == (n, 0) and errors == int((pred != obs).any(dim=1).sum().item())
    osd.close(); demh.close()


def test_bb144_bp_and_osd_equal_the_oracle_shot_by_shot():
    """BASELINE configs[4] at full size through the kernels big models get (compressed min-sum BP; OSD-0 with the fully
    reduced basis on two words per lane): the same BP iteration counts and the same error vectors as
    oracle/bp_oracle.c::orc_bp_osd0 (kernel schedule), bit for bit, on a sample of shots at p = 3e-3 -- where almost every
    shot goes through OSD."""
    code = gq.BivariateBicycleCode(depolarize1_rate=0.003, depolarize2_rate=0.003)
    code.build_memory_circuit(number_of_rounds=12, logic_check="Z")
    dem = code.memory_circuit.detector_error_model()
    demh = be.DemHandle(dem, device=0)
    osd = be.DecoderHandle(demh, be.BP_MINSUM, bp_max_iter=30, bp_alpha=0.9, bp_osd=True)
    orc = po.BpOracle(dem, max_iter=30, alpha=0.9)
    n = 48
    dets, obs = demh.sample(13, 0, n)
    pred, it_gpu, flags, corr = osd.decode(dets, want_weights=True, want_flags=True, want_corrections=True)
    bits = np.unpackbits(b8(dets, dem.num_detectors), axis=1, bitorder="little")[:, : dem.num_detectors]
    cb = np.unpackbits(corr.cpu().numpy().view(np.uint8).reshape(n, -1), axis=1, bitorder="little")[:, : dem.num_errors]
    it_gpu, flags = it_gpu.cpu().numpy(), flags.cpu().numpy()
    n_osd = 0
    for s in range(n):
        e, it, st, nb, used = orc.decode_osd(bits[s], max_candidates=-2048)
        assert it == it_gpu[s], (s, it, it_gpu[s])
        assert st != 1 and flags[s] == 0
        assert (e == cb[s]).all(), (s, it, st, nb, used)
        n_osd += st == 0
    assert n_osd >= n // 2
    osd.close(); demh.close()


@pytest.mark.parametrize("kind,d,p,nshots", [("UnrotatedSurfaceCode", 15, 0.003, 160), ("RotatedSurfaceCode", 17, 0.005, 96)])
def test_large_codes_match_the_oracle(kind, d, p, nshots):
    """BASELINE configs[2] (unrotated d = 15, 15 rounds: 6300 detectors, ~270 defects per shot) and a d = 17 point of the
    reference's own example sweep (~330 defects): the wide classes of the first tier on natural syndromes give the
    oracle's minimum weight, shot by shot."""
    dem = make_dem(kind, p, distance=d)
    demh = be.DemHandle(dem, device=0)
    dech = be.DecoderHandle(demh, be.MATCHING)
    dets, obs = demh.sample(3, 0, nshots)
    pred, wts, flags = dech.decode(dets, want_weights=True, want_flags=True)
    orc, _ = oracle_for(dem, dech)
    d8 = b8(dets, dem.num_detectors)
    nd = np.unpackbits(d8, axis=1, bitorder="little")[:, : dem.num_detectors].sum(1)
    assert nd.mean() > 200 and nd.max() > 256                 # beyond the 8-bit classes
    opred, owts = orc.decode_batch(d8, nthreads=8, return_weights=True)
    assert (flags.cpu().numpy() == 0).all() and (wts.cpu().numpy() == owts).all()
    assert (dech.decode(dets) == pred).all()
    dech.close(); demh.close()


def test_decode_dm_equals_decode_sm():
    """gq_decode_dm (detector-major bit-sliced in / out, the layout of gq_extract) == gq_decode on the same shots"""
    dem, demh, dech = case("rot5")
    n = 3 * demh.tile_shots + 37
    dets, obs = demh.sample(55, 0, n)
    pred = dech.decode(dets)
    dets_dm = demh.sm_to_dm(dets, demh.D, n)
    pred_dm = dech.decode_dm(dets_dm, n)
    assert (demh.dm_to_sm(pred_dm, 1, n) == pred).all()
    # SURVEY 8b "corr_optional": the same corrections as the shot-major entry point
    pred2, _, _, corr = dech.decode(dets, want_weights=True, want_flags=True, want_corrections=True)
    pred_dm2, corr_dm = dech.decode_dm(dets_dm, n, want_corrections=True)
    assert (pred_dm2 == pred_dm).all() and (corr_dm == corr).all()
    # more shots than one 2^20-shot slice, ragged: the slices are gathered / scattered on word boundaries
    dem, demh, dech = case("rep3")
    n = (1 << 20) + 5 * demh.tile_shots + 19
    dets, obs = demh.sample(56, 0, n)
    pred_dm = dech.decode_dm(demh.sm_to_dm(dets, demh.D, n), n)
    assert (demh.dm_to_sm(pred_dm, 1, n) == dech.decode(dets)).all()
