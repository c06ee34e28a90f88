"""Pin the CPU oracle itself (networkx, brute force) and the decoder's matching core compiled for the host."""
import ctypes
import itertools
import os
import subprocess

import networkx as nx
import numpy as np
import pytest

from graphqec_b200 import RepetitionCode, RotatedSurfaceCode
from oracle import pyoracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_blossom_equals_networkx_on_random_graphs():
    rng = np.random.default_rng(11)
    for _ in range(150):
        n = int(rng.integers(2, 13))
        w = np.triu(rng.integers(1, 25, size=(n, n)), 1)
        w = w * np.triu(rng.random((n, n)) < 0.7, 1)
        w = w + w.T
        tot, mate = po.max_weight_matching(w)
        g = nx.Graph()
        g.add_nodes_from(range(n))
        g.add_weighted_edges_from((i, j, int(w[i, j])) for i in range(n) for j in range(i + 1, n) if w[i, j])
        ref = sum(w[a, b] for a, b in nx.max_weight_matching(g))
        assert tot == ref
        for i in range(n):
            assert mate[i] == -1 or (mate[mate[i]] == i and w[i, mate[i]] > 0)


def _brute_force_min_weight(orc, defects):
    """exhaustive minimum over all pairings/boundary assignments (tiny cases only)"""
    dist = orc.dist_table().astype(np.int64)
    D = orc.nd
    best = None

    def rec(rest, acc):
        nonlocal best
        if best is not None and acc >= best:
            return
        if not rest:
            best = acc
            return
        a = rest[0]
        rec(rest[1:], acc + dist[a, D])
        for k in range(1, len(rest)):
            rec(rest[1:k] + rest[k + 1:], acc + dist[a, rest[k]])

    rec(list(defects), 0)
    return best


def test_oracle_mwpm_equals_brute_force():
    code = RotatedSurfaceCode(distance=3, depolarize1_rate=0.02, depolarize2_rate=0.02)
    code.build_memory_circuit(number_of_rounds=3)
    dem = code.memory_circuit.detector_error_model(decompose_errors=True)
    orc = po.MwpmOracle(dem)
    dets, obs = po.sample(dem, 5, 300)
    pred, wts = orc.decode_batch(dets, return_weights=True)
    bits = np.unpackbits(dets, axis=1, bitorder="little")[:, : dem.num_detectors]
    checked = 0
    for s in range(300):
        d = np.flatnonzero(bits[s]).tolist()
        if 0 < len(d) <= 8:
            assert wts[s] == _brute_force_min_weight(orc, d)
            checked += 1
    assert checked > 100


def test_oracle_sampler_rates_and_extract():
    rep = RepetitionCode(distance=5, depolarize1_rate=0.1, depolarize2_rate=0.1)
    rep.build_memory_circuit(number_of_rounds=5)
    dem = rep.memory_circuit.detector_error_model()
    n = 200000
    dets, obs = po.sample(dem, 3, n)
    H, L = dem.dense_check_matrices()
    # detector fire rate = (1 - prod(1-2p))/2 over the mechanisms touching it
    want = 0.5 * (1 - np.prod(np.where(H == 1, 1 - 2 * dem.probs[None, :], 1.0), axis=1))
    got = np.unpackbits(dets, axis=1, bitorder="little")[:, : dem.num_detectors].mean(axis=0)
    assert np.all(np.abs(got - want) < 5 * np.sqrt(want * (1 - want) / n))
    rng = np.random.default_rng(0)
    errs = (rng.random((64, dem.num_errors)) < 0.1).astype(np.uint8)
    d2, o2 = po.extract(dem, errs)
    assert (np.unpackbits(d2, axis=1, bitorder="little")[:, : dem.num_detectors] == (errs @ H.T) % 2).all()
    assert (np.unpackbits(o2, axis=1, bitorder="little")[:, : dem.num_observables] == (errs @ L.T) % 2).all()


def test_oracle_bp_decodes_single_errors():
    code = RotatedSurfaceCode(distance=3, depolarize1_rate=0.01, depolarize2_rate=0.01)
    code.build_memory_circuit(number_of_rounds=3)
    dem = code.memory_circuit.detector_error_model()
    bp = po.BpOracle(dem, max_iter=30, alpha=0.9)
    H, L = dem.dense_check_matrices()
    ok = 0
    for m in range(0, dem.num_errors, 3):
        e, it = bp.decode(H[:, m])
        if it > 0:
            assert ((H @ e) % 2 == H[:, m]).all()
            ok += 1
    # plain BP does not always converge on circuit-level DEMs (short cycles / degeneracy; this is
    # what OSD is for in the reference's bposd) -- most single faults must, though
    assert ok > 0.7 * len(range(0, dem.num_errors, 3))


@pytest.fixture(scope="module")
def hostsim():
    """blossom_warp.cuh compiled for the host with one lane (logic check only, never shipped)."""
    so = os.path.join(ROOT, "tests", "hostsim", "libhostsim.so")
    src = os.path.join(ROOT, "tests", "hostsim", "blossom_hostsim.cpp")
    hdr = os.path.join(ROOT, "graphqec_b200", "csrc", "blossom_warp.cuh")
    if not os.path.exists(so) or max(os.path.getmtime(src), os.path.getmtime(hdr)) > os.path.getmtime(so):
        subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-o", so, src])
    lib = ctypes.CDLL(so)
    lib.hostsim_solve.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    return lib


def test_decoder_matching_core_is_exact_on_host(hostsim):
    rng = np.random.default_rng(7)
    for _ in range(600):
        n = int(rng.integers(1, 14)) * 2
        hi = int(rng.choice([2, 3, 10, 1000]))
        w = np.triu(rng.integers(1, hi + 1, size=(n, n)), 1)
        w = (w + w.T).astype(np.uint16)
        mate = np.zeros(n, dtype=np.int16)
        wt = ctypes.c_longlong(0)
        rc = hostsim.hostsim_solve(n, w.ctypes.data, mate.ctypes.data, ctypes.byref(wt))
        assert rc == 0
        assert all(mate[mate[i]] == i and mate[i] != i for i in range(n))
        big = (int(w.max()) + 1) * (n + 2)
        tot, m2 = po.max_weight_matching(np.where(np.eye(n, dtype=bool), 0, big - w.astype(np.int64)))
        assert wt.value == sum(int(w[i, m2[i]]) for i in range(n) if m2[i] > i)
    c = (ctypes.c_longlong * 16)()
    hostsim.hostsim_counters(c)
    assert c[0] > 100 and c[1] > 10 and c[2] > 10  # shrink, expand and nested blossoms were exercised
    assert c[4] == 0                               # slack parity invariant never broken


def test_decoder_matching_core_handles_missing_edges(hostsim):
    w = np.full((4, 4), 0xFFFF, dtype=np.uint16)
    w[0, 1] = w[1, 0] = 5
    mate = np.zeros(4, dtype=np.int16)
    wt = ctypes.c_longlong(0)
    assert hostsim.hostsim_solve(4, w.ctypes.data, mate.ctypes.data, ctypes.byref(wt)) == 1


# ------------------------------------------------------------------------------------------
# first tier of the CUDA matching decoder (match_tree.cuh) compiled for the host with one lane
# ------------------------------------------------------------------------------------------

@pytest.fixture(scope="module", params=[128, 512, 1024], ids=["cap128_8bit_keys", "cap512_9bit_keys", "cap1024_10bit_keys"])
def tree_hostsim(request):
    """one-lane host build of match_tree.cuh; capacity 128 uses the 8-bit vertex field of the K <= 8 kernels,
    capacity 512 the 9-bit field of the K = 10, 12, 16 kernels, capacity 1024 the 10-bit field of K = 24, 32"""
    cap = request.param
    so = os.path.join(ROOT, "tests", "hostsim", f"libmthostsim_{cap}.so")
    srcs = [os.path.join(ROOT, "tests", "hostsim", f) for f in ("mt_hostsim.cpp", "warp_sim.h")]
    srcs += [os.path.join(ROOT, "graphqec_b200", "csrc", f) for f in ("match_tree.cuh", "blossom_regs.cuh")]
    if not os.path.exists(so) or max(os.path.getmtime(s) for s in srcs) > os.path.getmtime(so):
        subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", f"-DHS_CAP={cap}", "-DGQ_FP2_MAXK=1000000", "-o", so, srcs[0]])   # the one-lane build has K = cap: keep the second-event fast path in
    lib = ctypes.CDLL(so)
    lib.hostsim_mt_decode.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                      ctypes.c_long] + [ctypes.c_void_p] * 5
    lib.hostsim_mt_stats.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.cap = cap
    return lib


def _device_tables(orc):
    """the (D+2)x(D+2) distance / path-observable tables in the layout matching.cu uploads"""
    T, P = orc.dist_table(), orc.parity_table()
    D = orc.nd
    far = T >= 0xFFFF
    dist = np.full((D + 2, D + 2), 0xFFFF, np.uint16)
    pobs = np.zeros((D + 2, D + 2), np.uint8)
    Tc = np.where(far | (T < 0), 0xFFFF, T).astype(np.uint16)
    dist[:D, :D] = Tc[:, :D]; dist[:D, D] = Tc[:, D]; dist[D, :D] = Tc[:, D]
    pobs[:D, :D] = P[:, :D]; pobs[:D, D] = P[:, D]; pobs[D, :D] = P[:, D]
    np.fill_diagonal(dist, 0xFFFF)
    return dist, pobs


def _nbr_len():
    import re
    src = open(os.path.join(ROOT, "graphqec_b200", "csrc", "match_tree.cuh")).read()
    return int(re.search(r"#define GQ_NBR_LEN (\d+)", src).group(1))


def _neighbor_lists(dist, D, DW, L=None):
    L = L or _nbr_len()
    """nearest-detector lists as matching.cu builds them: distance << 16 | detector, ascending, padded"""
    nbr = np.full((D + 2, L), (0xFFFF << 16) | (32 * DW), np.uint32)
    for s in range(D):
        d = dist[s, :D].astype(np.uint32)
        ent = np.sort(((d << 16) | np.arange(D, dtype=np.uint32))[d != 0xFFFF])[:L]
        nbr[s, :len(ent)] = ent
    return nbr


def _tree_decode(lib, orc, dets_b8, lists=True):
    D = orc.nd
    dist, pobs = _device_tables(orc)
    nbr = _neighbor_lists(dist, D, (D + 31) // 32) if lists else None
    n = dets_b8.shape[0]
    DW = (D + 31) // 32
    rows = np.zeros((n, DW * 4), np.uint8)
    rows[:, :dets_b8.shape[1]] = dets_b8
    rows32 = np.ascontiguousarray(rows).view(np.uint32)
    pred = np.zeros(n, np.uint32); wt = np.zeros(n, np.int32); rc = np.zeros(n, np.int32)
    lib.hostsim_mt_decode(D, dist.ctypes.data, pobs.ctypes.data, DW, rows32.ctypes.data, n,
                          pred.ctypes.data, wt.ctypes.data, rc.ctypes.data, None,
                          nbr.ctypes.data if lists else None)
    return pred, wt, rc


def test_tree_tier_is_exact_on_surface_code_shots(tree_hostsim):
    code = RotatedSurfaceCode(distance=5, depolarize1_rate=0.01, depolarize2_rate=0.01)
    code.build_memory_circuit(number_of_rounds=5, logic_check="Z")
    dem = code.memory_circuit.detector_error_model(decompose_errors=True)
    orc = po.MwpmOracle(dem)
    dets, _ = po.sample(dem, 3, 3000)
    opred, owt = orc.decode_batch(dets, return_weights=True)
    for lists in (True, False):   # nearest-detector lists / every defect through the row fallback
        pred, wt, rc = _tree_decode(tree_hostsim, orc, dets, lists=lists)
        assert (rc == 0).all()
        assert (wt == owt).all()          # every matching is a minimum-weight one
        assert (pred != opred).mean() < 0.01   # equal-weight optima may differ in the predicted observable


def test_tree_tier_is_exact_on_random_graphs_with_blossoms(tree_hostsim):
    rng = np.random.default_rng(5)
    stats = (ctypes.c_longlong * 16)()
    tree_hostsim.hostsim_mt_stats(stats, 1)
    for trial in range(60):
        D = int(rng.integers(6, 40))
        edges = []
        for u in range(D):
            for v in range(u + 1, D):
                if v == u + 1 or rng.random() < 0.25:     # the chain keeps the graph connected
                    edges.append((u, v, 0.1, int(rng.integers(0, 2))))
        for u in range(D):
            if u == 0 or rng.random() < (0.3 if trial % 3 else 0.05):
                edges.append((u, -1, 0.1, int(rng.integers(0, 2))))
        hi = int(rng.choice([3, 10, 1000]))
        w = rng.integers(1, hi + 1, size=len(edges))
        orc = po.MwpmOracle(edges=edges, num_detectors=D, num_observables=1, int_weights=w)
        bits = (rng.random((200, D)) < rng.choice([0.15, 0.4, 0.8])).astype(np.uint8)
        dets = np.packbits(bits, axis=1, bitorder="little")
        pred, wt, rc = _tree_decode(tree_hostsim, orc, dets, lists=bool(trial % 2))
        opred, owt = orc.decode_batch(dets, return_weights=True)
        assert (rc == 0).all() and (wt == owt).all(), (trial, np.where(wt != owt)[0][:5])
    tree_hostsim.hostsim_mt_stats(stats, 0)
    assert stats[3] > 200 and stats[7] > 5    # shrinks and expands were exercised


def test_tree_tier_is_exact_on_dense_syndromes(tree_hostsim):
    """rotated d = 11 at p = 1.5 %: ~230 defects per shot; shots beyond the build's capacity report code 4"""
    code = RotatedSurfaceCode(distance=11, depolarize1_rate=0.015, depolarize2_rate=0.015)
    code.build_memory_circuit(number_of_rounds=11, logic_check="Z")
    dem = code.memory_circuit.detector_error_model(decompose_errors=True)
    orc = po.MwpmOracle(dem)
    dets, _ = po.sample(dem, 11, 200)
    nd = np.unpackbits(dets, axis=1, bitorder="little")[:, :dem.num_detectors].sum(1)
    opred, owt = orc.decode_batch(dets, return_weights=True)
    pred, wt, rc = _tree_decode(tree_hostsim, orc, dets)
    fits = nd <= tree_hostsim.cap
    assert (rc[fits] == 0).all() and (rc[~fits] == 4).all()
    assert (wt[fits] == owt[fits]).all()
    if tree_hostsim.cap >= 512:
        assert fits.all() and (nd > 128).sum() > 100


def test_tree_tier_reports_undecodable_syndromes(tree_hostsim):
    # detector 2 has no edge at all: a defect there cannot be matched
    orc = po.MwpmOracle(edges=[(0, 1, 0.1, 0), (0, -1, 0.1, 1)], num_detectors=3, num_observables=1,
                        int_weights=np.array([5, 7]))
    dets = np.packbits(np.array([[1, 1, 0], [0, 0, 1], [1, 0, 1]], np.uint8), axis=1, bitorder="little")
    pred, wt, rc = _tree_decode(tree_hostsim, orc, dets)
    assert rc.tolist() == [0, 1, 1] and wt[0] == 5


# ------------------------------------------------------------------------------------------
# OSD-0 (oracle/bp_oracle.c::orc_osd0) against the textbook definition
# ------------------------------------------------------------------------------------------

def _osd0_textbook(H, post, syn):
    """full OSD-0: greedy basis of ALL of H's column space in posterior order, then solve H_S x = s"""
    D, M = H.shape
    order = sorted(range(M), key=lambda v: (float(post[v]), v))
    rows = []       # reduced basis vectors as python ints: (vector, coefficient mask over chosen ordinals)
    piv = {}
    chosen = []
    for c in order:
        v = int("".join(str(int(b)) for b in H[::-1, c]), 2)
        m = 0
        while v:
            r = (v & -v).bit_length() - 1
            if r in piv:
                bv, bm = rows[piv[r]]
                v ^= bv; m ^= bm
            else:
                piv[r] = len(rows)
                rows.append((v, m | (1 << len(chosen))))
                chosen.append(c)
                break
    s = int("".join(str(int(b)) for b in syn[::-1]), 2)
    m = 0
    while s:
        r = (s & -s).bit_length() - 1
        if r not in piv:
            return None
        bv, bm = rows[piv[r]]
        s ^= bv; m ^= bm
    e = np.zeros(M, np.uint8)
    for b, c in enumerate(chosen):
        if (m >> b) & 1:
            e[c] = 1
    return e


def test_osd0_oracle_equals_the_textbook_definition():
    from graphqec_b200 import CssCode
    Hs = np.array([[1, 1, 1, 1, 0, 0, 0], [0, 1, 1, 0, 1, 1, 0], [0, 0, 1, 1, 0, 1, 1]])
    code = CssCode(Hx=Hs, Hz=Hs, depolarize1_rate=0.01, depolarize2_rate=0.01)
    code.build_memory_circuit(number_of_rounds=3, logic_check="Z")
    dem = code.memory_circuit.detector_error_model()
    H, L = dem.dense_check_matrices()
    bp = po.BpOracle(dem, max_iter=20)
    rng = np.random.default_rng(7)
    solved = 0
    for t in range(120):
        e = (rng.random(dem.num_errors) < 0.02).astype(np.uint8)
        syn = ((H @ e) % 2).astype(np.uint8)
        post = (bp.prior + rng.normal(0, 3.0, dem.num_errors)).astype(np.float32)    # any posterior will do
        if t % 3 == 0:
            post = np.round(post)                                                     # plenty of ties
        rc, got = bp.osd0(post, syn)
        want = _osd0_textbook(H, post, syn)
        assert (rc == 0) == (want is not None)
        if want is not None:
            solved += 1
            assert (got == want).all() and (((H @ got) % 2) == syn).all()
        # the bounded walk either gives the same answer or gives up
        rc2, got2 = bp.osd0(post, syn, max_candidates=40, max_basis=16)
        assert rc2 == 1 or (got2 == want).all()
    assert solved > 100


def _osd_kernel_schedule(H, post, syn, prefix):
    """the CUDA kernel's walk (orc_osd0 with max_candidates = -prefix), restated with python ints: reliability-ordered
    prefix, then the variables of the lowest detector left of the syndrome, then every variable"""
    D, M = H.shape
    cols = [int("".join(str(int(b)) for b in H[::-1, c]), 2) for c in range(M)]
    order = sorted(range(M), key=lambda v: (float(post[v]), v))[:prefix]
    rows, piv, chosen = [], {}, []
    s = int("".join(str(int(b)) for b in syn[::-1]), 2)
    sm = 0
    used = 0

    def reduce_s():
        nonlocal s, sm
        while s:
            r = (s & -s).bit_length() - 1
            if r not in piv:
                return r
            s ^= rows[piv[r]][0]; sm ^= rows[piv[r]][1]
        return -1

    def offer(c):
        nonlocal used
        used += 1
        v, m = cols[c], 0
        while v:
            r = (v & -v).bit_length() - 1
            if r in piv:
                v ^= rows[piv[r]][0]; m ^= rows[piv[r]][1]
            else:
                piv[r] = len(rows); rows.append((v, m | (1 << len(chosen)))); chosen.append(c)
                return

    rs = reduce_s()
    for c in order:
        if rs < 0:
            break
        offer(c); rs = reduce_s()
    tr, k, stage2 = -1, 0, False
    while rs >= 0 and not stage2:
        if rs != tr:
            tr, k = rs, 0
            row = [c for c in range(M) if H[tr, c]]
        if k >= len(row):
            stage2 = True
            break
        offer(row[k]); k += 1; rs = reduce_s()
    if stage2:
        for c in range(M):
            if rs < 0:
                break
            offer(c); rs = reduce_s()
    if rs >= 0:
        return None, used
    e = np.zeros(M, np.uint8)
    for b, c in enumerate(chosen):
        if (sm >> b) & 1:
            e[c] = 1
    return e, used


def test_osd_kernel_schedule_targets_the_detectors_left_of_the_syndrome():
    """max_candidates < 0 (the CUDA kernel's walk): when the ordered prefix does not span the syndrome the walk goes on
    through the variables of the lowest detector still set -- identical to the python restatement above, always a valid
    correction when one exists, and far fewer candidates than a blind walk through all variables."""
    from graphqec_b200 import CssCode
    Hs = np.array([[1, 1, 1, 1, 0, 0, 0], [0, 1, 1, 0, 1, 1, 0], [0, 0, 1, 1, 0, 1, 1]])
    code = CssCode(Hx=Hs, Hz=Hs, depolarize1_rate=0.01, depolarize2_rate=0.01)
    code.build_memory_circuit(number_of_rounds=4, logic_check="Z")
    dem = code.memory_circuit.detector_error_model()
    H, L = dem.dense_check_matrices()
    M = dem.num_errors
    bp = po.BpOracle(dem, max_iter=20)
    rng = np.random.default_rng(11)
    import ctypes
    tail_used = 0
    for t in range(80):
        e = (rng.random(M) < 0.03).astype(np.uint8)
        syn = ((H @ e) % 2).astype(np.uint8)
        post = rng.normal(0, 3.0, M).astype(np.float32)
        prefix = int(rng.integers(1, 12))
        out = np.zeros(M, np.uint8)
        nb, used = ctypes.c_int(0), ctypes.c_int(0)
        rc = po.lib().orc_osd0(dem.num_detectors, M, po._p(bp.chk_indptr), po._p(bp.chk_var), po._p(bp.var_indptr), po._p(bp.var_edge),
                               po._p(post), po._p(syn), -prefix, 0, po._p(out), ctypes.byref(nb), ctypes.byref(used))
        want, want_used = _osd_kernel_schedule(H, post, syn, prefix)
        assert (rc == 0) == (want is not None)
        if want is not None:
            assert (out == want).all() and (((H @ out) % 2) == syn).all() and used.value == want_used
            tail_used += used.value > prefix
            assert used.value < prefix + 8 * int(H.sum(axis=1).max()) + 1 or used.value > M       # targeted, or the full fallback
    assert tail_used > 40


def test_bp_then_osd_always_reproduces_the_syndrome():
    code = RotatedSurfaceCode(distance=3, depolarize1_rate=0.02, depolarize2_rate=0.02)
    code.build_memory_circuit(number_of_rounds=3, logic_check="Z")
    dem = code.memory_circuit.detector_error_model()
    H, L = dem.dense_check_matrices()
    bp = po.BpOracle(dem, max_iter=10)
    dets, _ = po.sample(dem, 5, 400)
    bits = np.unpackbits(dets, axis=1, bitorder="little")[:, : dem.num_detectors]
    used = 0
    for s in range(400):
        e, it, st, nb, cand = bp.decode_osd(bits[s])
        assert st in (-1, 0) and (((H @ e) % 2) == bits[s]).all()
        used += st == 0
    assert used > 10


# ------------------------------------------------------------------------------------------
# The two reformulations the big-model GPU kernels rely on, restated in python and held against the oracle on the CPU
# (the kernels themselves are compared with the oracle bit for bit in tests/test_gpu_parity.py)
# ------------------------------------------------------------------------------------------

def _compressed_min_sum(dem, syn, max_iter, alpha):
    """bp_cms_kernel's formulation: no per-edge message is stored -- per check (min1, argmin, min2, sign) and one sign bit
    per edge; a variable rebuilds its incoming messages from its checks' records, and offers its outgoing ones to the
    checks' next records.  fp32, the sums in ascending check order."""
    f32 = np.float32
    D, M = dem.num_detectors, dem.num_errors
    ptr, idx = dem.det_indptr.astype(int), dem.det_idx.astype(int)
    p = np.clip(dem.probs, 1e-30, 1 - 1e-7)
    prior = np.log((1 - p) / p).astype(f32)
    alpha = f32(alpha)
    m1o, m2o, argo, sgo = np.zeros(D, f32), np.zeros(D, f32), np.zeros(D, int), np.zeros(D, int)
    esign = {}
    for it in range(1, max_iter + 2):
        m1n, m2n = np.full(D, np.inf, f32), np.full(D, np.inf, f32)
        argn, sgn, par = np.full(D, -1, int), np.zeros(D, int), np.zeros(D, int)
        dec = np.zeros(M, np.uint8)
        for v in range(M):
            cs = sorted(idx[ptr[v]:ptr[v + 1]])
            tot, cv = prior[v], []
            for c in cs:
                mag = m2o[c] if argo[c] == v else m1o[c]
                cv.append(f32(alpha * (f32(-mag) if sgo[c] ^ esign.get((v, c), 0) else mag)))
            for x in cv:
                tot = f32(tot + x)
            bit = int(tot < 0)
            dec[v] = bit
            for c, x in zip(cs, cv):
                m = f32(tot - x)
                neg = int(m < 0)
                esign[(v, c)] = neg
                sgn[c] ^= neg
                par[c] ^= bit
                a = abs(m)
                if a < m1n[c] or (a == m1n[c] and v < argn[c]):
                    m2n[c] = m1n[c]; m1n[c] = a; argn[c] = v
                elif a < m2n[c]:
                    m2n[c] = a
        bad = (par != syn).any()
        m1o, m2o, argo, sgo = m1n, m2n, argn, sgn ^ syn
        if it > 1 and not bad:
            return dec, it - 1
    return dec, -max_iter


def test_compressed_min_sum_equals_the_flooding_schedule():
    code = RotatedSurfaceCode(distance=3, depolarize1_rate=0.01, depolarize2_rate=0.01)
    code.build_memory_circuit(number_of_rounds=3, logic_check="Z")
    dem = code.memory_circuit.detector_error_model()
    H, L = dem.dense_check_matrices()
    orc = po.BpOracle(dem, max_iter=12, alpha=0.9)
    rng = np.random.default_rng(5)
    iters = set()
    for s in range(40):
        e = (rng.random(dem.num_errors) < np.clip(dem.probs * 3, 0, 0.5)).astype(np.uint8)
        syn = ((H @ e) % 2).astype(np.uint8)
        want, it = orc.decode(syn)
        got, it2 = _compressed_min_sum(dem, syn, 12, 0.9)
        assert it == it2 and (want == got).all()
        iters.add(it)
    assert len(iters) >= 4 and min(iters) < 0          # converged after different numbers of iterations, and not at all


def _osd_fully_reduced(H, post, syn, prefix):
    """osd_rref_kernel's formulation of the kernel schedule: the basis is kept zero at all other pivots, a column is
    reduced in one step by the vectors of the pivots among its own detectors, compositions are kept in basis columns."""
    D, M = H.shape
    cols = [int("".join(str(int(b)) for b in H[::-1, c]), 2) for c in range(M)]
    order = sorted(range(M), key=lambda v: (float(post[v]), v))[:prefix]
    rows, coefs, piv, chosen = [], [], {}, []
    state = {"s": int("".join(str(int(b)) for b in syn[::-1]), 2), "sc": 0, "used": 0}
    low = lambda x: (x & -x).bit_length() - 1 if x else -1

    def offer(c):
        state["used"] += 1
        v, vc, h = cols[c], 0, cols[c]
        while h:
            d = low(h); h &= h - 1
            if d in piv:
                v ^= rows[piv[d]]; vc ^= coefs[piv[d]]
        assert all(((v >> d) & 1) == 0 for d in piv)          # one step suffices
        if v == 0:
            return
        r, nb = low(v), len(rows)
        vc |= 1 << nb
        for i in range(nb):
            if (rows[i] >> r) & 1:
                rows[i] ^= v; coefs[i] ^= vc
        rows.append(v); coefs.append(vc); piv[r] = nb; chosen.append(c)
        if (state["s"] >> r) & 1:
            state["s"] ^= v; state["sc"] ^= vc

    for c in order:
        if state["s"] == 0:
            break
        offer(c)
    tr, k, row, stage2 = -1, 0, [], False
    while state["s"]:
        rs = low(state["s"])
        if rs != tr:
            tr, k, row = rs, 0, [c for c in range(M) if H[rs, c]]
        if k >= len(row):
            stage2 = True
            break
        offer(row[k]); k += 1
    if stage2:
        for c in range(M):
            if state["s"] == 0:
                break
            offer(c)
    if state["s"]:
        return None, state["used"]
    e = np.zeros(M, np.uint8)
    for b, c in enumerate(chosen):
        if (state["sc"] >> b) & 1:
            e[c] = 1
    return e, state["used"]


def test_fully_reduced_osd_equals_the_echelon_walk():
    """same candidates, same independence decisions, same detector targeted by the walk past the prefix (the lowest bit
    left of the syndrome does not depend on the form of the basis), same unique solution"""
    import ctypes
    from graphqec_b200 import CssCode
    Hs = np.array([[1, 1, 1, 1, 0, 0, 0], [0, 1, 1, 0, 1, 1, 0], [0, 0, 1, 1, 0, 1, 1]])
    code = CssCode(Hx=Hs, Hz=Hs, depolarize1_rate=0.01, depolarize2_rate=0.01)
    code.build_memory_circuit(number_of_rounds=4, logic_check="Z")
    dem = code.memory_circuit.detector_error_model()
    H, L = dem.dense_check_matrices()
    M = dem.num_errors
    bp = po.BpOracle(dem, max_iter=20)
    rng = np.random.default_rng(3)
    tails = 0
    for t in range(120):
        e = (rng.random(M) < 0.03).astype(np.uint8)
        syn = ((H @ e) % 2).astype(np.uint8)
        post = rng.normal(0, 3.0, M).astype(np.float32)
        prefix = int(rng.integers(1, 40))
        out = np.zeros(M, np.uint8)
        nb, used = ctypes.c_int(0), ctypes.c_int(0)
        rc = po.lib().orc_osd0(dem.num_detectors, M, po._p(bp.chk_indptr), po._p(bp.chk_var), po._p(bp.var_indptr), po._p(bp.var_edge),
                               po._p(post), po._p(syn), -prefix, 0, po._p(out), ctypes.byref(nb), ctypes.byref(used))
        want, want_used = _osd_fully_reduced(H, post, syn, prefix)
        assert (rc == 0) == (want is not None)
        if want is not None:
            assert (out == want).all() and used.value == want_used
        tails += used.value > prefix
    assert tails > 60


def test_jump_start_reaches_the_same_optimum():
    """The blossom port's optional greedy start (used only for the timed CPU baseline) changes the schedule, not the
    answer: same total weight on random dense graphs (ties included) and on surface-code shots, same failure count."""
    rng = np.random.default_rng(1)
    L = po.lib()
    try:
        for t in range(150):
            n = 2 * int(rng.integers(1, 20))
            w = rng.integers(1, 4 if t % 3 == 0 else 60, size=(n, n)).astype(np.int64)
            w = np.triu(w, 1); w = w + w.T
            ww = np.where(np.eye(n, dtype=bool), 0, (w.max() + 1) * (n + 2) - w).astype(np.int64)
            po.set_jump_start(False); t0, m0 = po.max_weight_matching(ww)
            po.set_jump_start(True); t1, m1 = po.max_weight_matching(ww)
            assert t0 == t1 and (np.asarray(m1) >= 0).all() and all(m1[m1[i]] == i for i in range(n))
        code = RotatedSurfaceCode(distance=5, depolarize1_rate=0.008, depolarize2_rate=0.008)
        code.build_memory_circuit(number_of_rounds=5, logic_check="Z")
        dem = code.memory_circuit.detector_error_model(decompose_errors=True)
        orc = po.MwpmOracle(dem)
        dets, obs = po.sample(dem, 21, 3000)
        po.set_jump_start(False); p0, w0 = orc.decode_batch(dets, nthreads=2, return_weights=True)
        po.set_jump_start(True); p1, w1 = orc.decode_batch(dets, nthreads=2, return_weights=True)
        assert (w0 == w1).all() and w0.max() > 0
        assert (p0 != p1).sum() <= 3                                  # equal-weight optima of different parity are rare
    finally:
        po.set_jump_start(False)
