"""ORACLE -- TEST INFRASTRUCTURE ONLY.

ctypes front-end of ``oracle/liboracle.so`` plus the pure-Python pieces of the CPU restatement of
the reference's hot path.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this; the product package
``graphqec_b200`` never does.  PARITY UNPINNED against the reference's wheels (Stim, sinter,
PyMatching are not installed; see the header of ``mwpm_oracle.c``).

Follows, for the matching graph, PyMatching's conventions as restated in SURVEY.md section 8a row
A8: every graphlike component of every mechanism becomes an edge (two detectors) or a boundary
edge (one detector) carrying the mechanism's probability; parallel edges merge with
p <- p1(1-p2) + p2(1-p1); weight = ln((1-p)/p).
"""

from __future__ import annotations

import ctypes
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

NO_DET = 0xFFFFFFFF
WEIGHT_LEVELS = 1000  # integer weight of the heaviest edge (shared convention with the CUDA decoder)


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("mwpm_oracle.c", "sampler_oracle.c", "bp_oracle.c")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        c = ctypes
        L.orc_max_weight_matching.restype = c.c_longlong
        L.orc_max_weight_matching.argtypes = [c.c_int, c.c_void_p, c.c_void_p]
        L.orc_set_jump_start.argtypes = [c.c_int]
        L.orc_graph_create.restype = c.c_void_p
        L.orc_graph_create.argtypes = [c.c_int, c.c_int, c.c_int] + [c.c_void_p] * 4
        L.orc_graph_free.argtypes = [c.c_void_p]
        L.orc_graph_dist.restype = c.c_void_p
        L.orc_graph_dist.argtypes = [c.c_void_p]
        L.orc_graph_par.restype = c.c_void_p
        L.orc_graph_par.argtypes = [c.c_void_p]
        L.orc_decode_batch.restype = c.c_long
        L.orc_decode_batch.argtypes = [c.c_void_p, c.c_void_p, c.c_long, c.c_int, c.c_void_p, c.c_void_p, c.c_int]
        L.orc_decode_one.restype = c.c_int
        L.orc_decode_one.argtypes = [c.c_void_p, c.c_void_p, c.c_int, c.c_void_p, c.c_void_p]
        L.orc_extract.argtypes = [c.c_int] * 3 + [c.c_void_p] * 5 + [c.c_long, c.c_void_p, c.c_void_p]
        L.orc_sample.argtypes = [c.c_int] * 3 + [c.c_void_p] * 5 + [c.c_uint64, c.c_long, c.c_void_p, c.c_void_p]
        L.orc_count_errors.restype = c.c_long
        L.orc_count_errors.argtypes = [c.c_void_p, c.c_void_p, c.c_long, c.c_int]
        L.orc_bp_minsum.restype = c.c_int
        L.orc_bp_minsum.argtypes = [c.c_int, c.c_int] + [c.c_void_p] * 6 + [c.c_int, c.c_float, c.c_void_p]
        L.orc_bp_osd0.restype = c.c_int
        L.orc_bp_osd0.argtypes = [c.c_int, c.c_int] + [c.c_void_p] * 6 + [c.c_int, c.c_float, c.c_int, c.c_int] + [c.c_void_p] * 4
        L.orc_osd0.restype = c.c_int
        L.orc_osd0.argtypes = [c.c_int, c.c_int] + [c.c_void_p] * 6 + [c.c_int, c.c_int] + [c.c_void_p] * 3
        _LIB = L
    return _LIB


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


# ------------------------------------------------------------------------------------------
# matching graph (A8)
# ------------------------------------------------------------------------------------------

def merged_edges(dem) -> list[tuple[int, int, float, int]]:
    """``(u, v or -1, probability, observable mask)`` after merging parallel edges; u < v."""
    a, b, om, p = dem.graphlike_components()
    table: dict[tuple[int, int], list] = {}
    for x, y, o, pp in zip(a.tolist(), b.tolist(), om.tolist(), p.tolist()):
        if y == NO_DET:
            key = (x, -1)
        else:
            key = (min(x, y), max(x, y))
        ent = table.get(key)
        if ent is None:
            table[key] = [pp, o]
        else:
            ent[0] = ent[0] * (1 - pp) + pp * (1 - ent[0])
            # PyMatching keeps the observable mask of the first edge it saw
    return [(k[0], k[1], v[0], v[1]) for k, v in sorted(table.items())]


def integer_weights(probs: np.ndarray, levels: int = WEIGHT_LEVELS) -> np.ndarray:
    """Discretised weights: the heaviest edge maps to ``levels``; every edge weighs at least 1."""
    probs = np.clip(np.asarray(probs, dtype=np.float64), 1e-300, 0.5)
    w = np.log((1.0 - probs) / probs)
    wmax = float(w.max()) if w.size else 1.0
    if wmax <= 0:
        return np.ones_like(w, dtype=np.int64)
    return np.maximum(1, np.rint(w * (levels / wmax))).astype(np.int64)


class MwpmOracle:
    """Exact MWPM decoder on the CPU (see ``mwpm_oracle.c``)."""

    def __init__(self, dem=None, edges=None, num_detectors=None, num_observables=None, int_weights=None):
        if edges is None:
            edges = merged_edges(dem)
            num_detectors, num_observables = dem.num_detectors, dem.num_observables
        self.edges = edges
        self.nd, self.no = int(num_detectors), int(num_observables)
        if self.no > 8:
            raise ValueError("oracle supports at most 8 observables")
        eu = np.array([e[0] for e in edges], dtype=np.int32)
        ev = np.array([e[1] for e in edges], dtype=np.int32)
        if int_weights is None:
            int_weights = integer_weights(np.array([e[2] for e in edges]))
        self.weights = np.asarray(int_weights, dtype=np.int64)
        eo = np.array([e[3] for e in edges], dtype=np.uint8)
        self._keep = (eu, ev, eo)
        self._h = lib().orc_graph_create(self.nd, self.no, len(edges), _p(eu), _p(ev), _p(self.weights), _p(eo))

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().orc_graph_free(self._h)
                self._h = None
        except Exception:   # interpreter shutdown
            pass

    def dist_table(self) -> np.ndarray:
        n = self.nd
        buf = (ctypes.c_int32 * (n * (n + 1))).from_address(lib().orc_graph_dist(self._h))
        return np.frombuffer(buf, dtype=np.int32).reshape(n, n + 1)

    def parity_table(self) -> np.ndarray:
        n = self.nd
        buf = (ctypes.c_uint8 * (n * (n + 1))).from_address(lib().orc_graph_par(self._h))
        return np.frombuffer(buf, dtype=np.uint8).reshape(n, n + 1)

    def decode_batch(self, dets_b8: np.ndarray, nthreads: int = 1, return_weights: bool = False):
        dets_b8 = np.ascontiguousarray(dets_b8, dtype=np.uint8)
        n = dets_b8.shape[0]
        pred = np.zeros(n, dtype=np.uint8)
        wts = np.zeros(n, dtype=np.int64) if return_weights else None
        nfail = lib().orc_decode_batch(
            self._h, _p(dets_b8), n, dets_b8.shape[1], _p(pred), _p(wts) if wts is not None else None, nthreads
        )
        self.last_failures = int(nfail)
        return (pred, wts) if return_weights else pred

    def decode_one(self, defects) -> tuple[int, list[tuple[int, int]], int]:
        d = np.asarray(defects, dtype=np.int32)
        pairs = np.full(2 * max(1, len(d)), -2, dtype=np.int32)
        wt = ctypes.c_longlong(0)
        pred = lib().orc_decode_one(self._h, _p(d), len(d), _p(pairs), ctypes.byref(wt))
        return pred, [(int(pairs[2 * i]), int(pairs[2 * i + 1])) for i in range(len(d))], int(wt.value)


def max_weight_matching(w: np.ndarray) -> tuple[int, np.ndarray]:
    w = np.ascontiguousarray(w, dtype=np.int64)
    n = w.shape[0]
    mate = np.zeros(max(n, 1), dtype=np.int32)
    tot = lib().orc_max_weight_matching(n, _p(w), _p(mate))
    return int(tot), mate[:n]


# ------------------------------------------------------------------------------------------
# sampler / extractor / counter (A7, A9)
# ------------------------------------------------------------------------------------------

def extract(dem, errs: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """Dense GF(2) H.e, L.e for explicit error vectors ``errs`` uint8[nshots, M] -> b8 rows."""
    errs = np.ascontiguousarray(errs, dtype=np.uint8)
    n = errs.shape[0]
    dets = np.zeros((n, (dem.num_detectors + 7) // 8), dtype=np.uint8)
    obs = np.zeros((n, max(1, (dem.num_observables + 7) // 8)), dtype=np.uint8)
    lib().orc_extract(
        dem.num_detectors, dem.num_observables, dem.num_errors,
        _p(dem.det_indptr), _p(dem.det_idx), _p(dem.obs_indptr), _p(dem.obs_idx),
        _p(errs), n, _p(dets), _p(obs),
    )
    return dets, obs


def sample(dem, seed: int, nshots: int) -> tuple[np.ndarray, np.ndarray]:
    dets = np.zeros((nshots, (dem.num_detectors + 7) // 8), dtype=np.uint8)
    obs = np.zeros((nshots, max(1, (dem.num_observables + 7) // 8)), dtype=np.uint8)
    probs = np.ascontiguousarray(dem.probs, dtype=np.float64)
    lib().orc_sample(
        dem.num_detectors, dem.num_observables, dem.num_errors, _p(probs),
        _p(dem.det_indptr), _p(dem.det_idx), _p(dem.obs_indptr), _p(dem.obs_idx),
        seed, nshots, _p(dets), _p(obs),
    )
    return dets, obs


def count_errors(pred: np.ndarray, obs: np.ndarray) -> int:
    pred = np.ascontiguousarray(pred, dtype=np.uint8).reshape(len(obs), -1)
    obs = np.ascontiguousarray(obs, dtype=np.uint8)
    return int(lib().orc_count_errors(_p(pred), _p(obs), len(obs), obs.shape[1]))


def set_jump_start(on: bool) -> None:
    """Process-wide switch of the blossom solver's greedy dual-feasible start (``mwpm_oracle.c::blossom_solve``): same
    optimum weight, ~4x fewer phases.  Off by default -- every parity check runs the plain textbook schedule; ``bench.py``
    turns it on for the timed CPU baseline so that the baseline is not handicapped."""
    lib().orc_set_jump_start(1 if on else 0)


def collect(dem, seed: int, nshots: int, nthreads: int = 1, oracle: MwpmOracle | None = None, chunk: int = 20000):
    """CPU restatement of the whole path: sample -> decode (exact MWPM) -> count."""
    import concurrent.futures as cf

    oracle = oracle or MwpmOracle(dem)
    errors = 0
    done = 0
    sizes = []
    while done < nshots:
        sizes.append(min(chunk, nshots - done))
        done += sizes[-1]
    # the (single-threaded) sampler of chunk k + 1 runs beside the (multi-threaded) decoder of chunk k: same chunks, same
    # seeds, same result as the plain loop -- the serial sampler just stops costing the multi-threaded baseline wall time
    with cf.ThreadPoolExecutor(1) as ex:
        nxt = ex.submit(sample, dem, seed, sizes[0]) if sizes else None
        for k, n in enumerate(sizes):
            dets, obs = nxt.result()
            nxt = ex.submit(sample, dem, seed + 7919 * (k + 1), sizes[k + 1]) if k + 1 < len(sizes) else None
            pred = oracle.decode_batch(dets, nthreads=nthreads)
            errors += count_errors(pred, obs)
    return done, errors


# ------------------------------------------------------------------------------------------
# min-sum BP (A10 restated without OSD)
# ------------------------------------------------------------------------------------------

class BpOracle:
    def __init__(self, dem, max_iter: int = 30, alpha: float = 0.9):
        self.dem = dem
        D, M = dem.num_detectors, dem.num_errors
        rows = dem.det_idx.astype(np.int64)
        cols = np.repeat(np.arange(M, dtype=np.int64), np.diff(dem.det_indptr.astype(np.int64)))
        order = np.lexsort((cols, rows))
        self.chk_var = cols[order].astype(np.uint32)
        self.chk_indptr = np.zeros(D + 1, dtype=np.uint32)
        np.add.at(self.chk_indptr, rows + 1, 1)
        self.chk_indptr = np.cumsum(self.chk_indptr).astype(np.uint32)
        # transpose map: for each variable, the check-major edge ids, ascending
        eid = np.arange(len(order), dtype=np.int64)
        by_var = np.lexsort((eid, self.chk_var.astype(np.int64)))
        self.var_edge = eid[by_var].astype(np.uint32)
        self.var_indptr = dem.det_indptr.astype(np.uint32)
        p = np.clip(dem.probs, 1e-30, 1 - 1e-7)
        self.prior = np.log((1 - p) / p).astype(np.float32)
        self.max_iter, self.alpha = max_iter, alpha

    def decode(self, syndrome_bits: np.ndarray) -> tuple[np.ndarray, int]:
        syn = np.ascontiguousarray(syndrome_bits, dtype=np.uint8)
        out = np.zeros(self.dem.num_errors, dtype=np.uint8)
        it = lib().orc_bp_minsum(
            self.dem.num_detectors, self.dem.num_errors, _p(self.chk_indptr), _p(self.chk_var),
            _p(self.var_indptr), _p(self.var_edge), _p(self.prior), _p(syn),
            self.max_iter, ctypes.c_float(self.alpha), _p(out),
        )
        return out, it

    def decode_osd(self, syndrome_bits: np.ndarray, max_candidates: int = 0, max_basis: int = 0):
        """BP, then OSD-0 where BP did not converge: (error vector, BP iterations (negative = not converged),
        OSD status (-1 not needed, 0 solved, 1 gave up), basis size, candidates walked)."""
        syn = np.ascontiguousarray(syndrome_bits, dtype=np.uint8)
        out = np.zeros(self.dem.num_errors, dtype=np.uint8)
        st, nb, used = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
        it = lib().orc_bp_osd0(
            self.dem.num_detectors, self.dem.num_errors, _p(self.chk_indptr), _p(self.chk_var),
            _p(self.var_indptr), _p(self.var_edge), _p(self.prior), _p(syn),
            self.max_iter, ctypes.c_float(self.alpha), int(max_candidates), int(max_basis), _p(out),
            ctypes.byref(st), ctypes.byref(nb), ctypes.byref(used),
        )
        return out, it, st.value, nb.value, used.value

    def osd0(self, posterior: np.ndarray, syndrome_bits: np.ndarray, max_candidates: int = 0, max_basis: int = 0):
        """OSD-0 alone on a given posterior: (status, error vector)."""
        syn = np.ascontiguousarray(syndrome_bits, dtype=np.uint8)
        post = np.ascontiguousarray(posterior, dtype=np.float32)
        out = np.zeros(self.dem.num_errors, dtype=np.uint8)
        rc = lib().orc_osd0(
            self.dem.num_detectors, self.dem.num_errors, _p(self.chk_indptr), _p(self.chk_var),
            _p(self.var_indptr), _p(self.var_edge), _p(post), _p(syn), int(max_candidates), int(max_basis), _p(out), None, None,
        )
        return rc, out


def wilson(errors: int, shots: int, z: float = 1.959963984540054) -> tuple[float, float]:
    if shots <= 0:
        return (0.0, 1.0)
    ph = errors / shots
    den = 1 + z * z / shots
    c = (ph + z * z / (2 * shots)) / den
    h = z * math.sqrt(ph * (1 - ph) / shots + z * z / (4 * shots * shots)) / den
    return (max(0.0, c - h), min(1.0, c + h))
