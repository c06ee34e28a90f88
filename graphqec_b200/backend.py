"""ctypes binding of ``libgqec_b200.so`` (C ABI: ``include/gqec_b200.h``).

PyTorch is used for device memory only (``tensor.data_ptr()``); all arithmetic of the hot path is in
the library's sm_100a kernels.  There is no CPU fallback: if the shared library is missing or no
CUDA device is present, every entry point raises.
"""

from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_double, c_float, c_int, c_int32, c_uint8, c_uint32, c_uint64, c_void_p

import numpy as np

__all__ = ["load_library", "library_path", "circuit_to_dem_arrays", "DemHandle", "DecoderHandle", "GqError", "MATCHING", "BP_MINSUM"]

MATCHING = 1
BP_MINSUM = 2
NO_DET = 0xFFFFFFFF

_LIB = None


class GqError(RuntimeError):
    pass


class _Params(ctypes.Structure):
    _fields_ = [
        ("weight_levels", c_int32),
        ("fast_capacity", c_int32),
        ("with_corrections", c_int32),
        ("bp_max_iter", c_int32),
        ("bp_alpha", c_float),
        ("legacy_tiers", c_int32),
        ("bp_osd", c_int32),
        ("bp_scratch", c_int32),
        ("reserved", c_int32 * 5),
    ]


class _DemInfo(ctypes.Structure):
    _fields_ = [
        ("num_detectors", c_uint32), ("num_observables", c_uint32), ("num_errors", c_uint32),
        ("det_words", c_uint32), ("obs_words", c_uint32), ("tile_shots", c_uint32),
        ("num_classes", c_uint32), ("num_segments", c_uint32), ("sum_p", c_double),
    ]


class _GraphInfo(ctypes.Structure):
    _fields_ = [
        ("num_nodes", c_uint32), ("num_edges", c_uint32), ("num_boundary_edges", c_uint32),
        ("fast_capacity", c_uint32), ("max_capacity", c_uint32), ("max_distance", c_uint32),
        ("reserved", c_uint32 * 6),
    ]


def library_path() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "libgqec_b200.so")


# name -> (restype, argtypes); also the list the symbol-export test checks against the header
SIGNATURES = {
    "gq_last_error": (ctypes.c_char_p, []),
    "gq_version": (c_int, []),
    "gq_device_count": (c_int, []),
    "gq_dem_create": (c_int, [c_uint32, c_uint32, c_uint32] + [c_void_p] * 9 + [c_int, POINTER(c_void_p)]),
    "gq_dem_destroy": (None, [c_void_p]),
    "gq_dem_get_info": (c_int, [c_void_p, POINTER(_DemInfo)]),
    "gq_sample": (c_int, [c_void_p, c_uint64, c_uint64, c_uint64, c_void_p, c_void_p, c_void_p]),
    "gq_extract": (c_int, [c_void_p, c_void_p, c_uint64, c_void_p, c_void_p]),
    "gq_transpose_dm_to_sm": (c_int, [c_void_p, c_void_p, c_uint32, c_uint64, c_void_p]),
    "gq_transpose_sm_to_dm": (c_int, [c_void_p, c_void_p, c_uint32, c_uint64, c_void_p]),
    "gq_decoder_create": (c_int, [c_void_p, c_int, POINTER(_Params), POINTER(c_void_p)]),
    "gq_decoder_destroy": (None, [c_void_p]),
    "gq_decoder_get_graph_info": (c_int, [c_void_p, POINTER(_GraphInfo)]),
    "gq_decoder_get_edges": (c_int, [c_void_p] * 6),
    "gq_decode": (c_int, [c_void_p, c_void_p, c_uint64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "gq_decode_b8": (c_int, [c_void_p, c_void_p, c_uint64, c_void_p]),
    "gq_decode_dm": (c_int, [c_void_p, c_void_p, c_uint64, c_void_p, c_void_p]),
    "gq_count_errors": (c_int, [c_void_p, c_void_p, c_void_p, c_uint64, c_void_p]),
    "gq_collect": (c_int, [c_void_p, c_void_p, c_uint64, c_uint64, c_uint64, c_uint64, POINTER(c_uint64 * 3)]),
    "gq_last_timing": (c_int, [c_void_p, POINTER(c_double * 8)]),
    "gq_last_kernel_profile": (c_int, [c_void_p, POINTER(c_double * 4)]),
    "gq_sync": (c_int, [c_void_p]),
    "gq_circuit_to_dem": (c_int, [ctypes.c_char_p, c_int, POINTER(c_void_p)]),
    "gq_dem_build_sizes": (c_int, [c_void_p, POINTER(c_uint32 * 6)]),
    "gq_dem_build_export": (c_int, [c_void_p] * 10),
    "gq_dem_build_destroy": (None, [c_void_p]),
}


def load_library() -> ctypes.CDLL:
    """Load the CUDA library; fails loudly when it has not been built (no fallback exists)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise GqError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C graphqec_b200/csrc`).  graphqec_b200 has no CPU fallback for the hot path."
        )
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def _check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load_library().gq_last_error()
        raise GqError(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")


def circuit_to_dem_arrays(stim_text: str, decompose_errors: bool):
    """Native circuit -> DEM (``gq_circuit_to_dem``, host code of the library, no device needed).

    Returns ``(D, O, probs, det_indptr, det_idx, obs_indptr, obs_idx, comps)`` with ``comps`` =
    ``(comp_indptr, comp_a, comp_b, comp_obs)`` or ``None``.  Raises ``ValueError`` with Stim's wording
    when a mechanism cannot be decomposed (callers fall back to the undecomposed model like sinter)."""
    lib = load_library()
    h = c_void_p()
    rc = lib.gq_circuit_to_dem(stim_text.encode(), 1 if decompose_errors else 0, ctypes.byref(h))
    if rc != 0:
        msg = lib.gq_last_error()
        raise ValueError(msg.decode() if msg else f"gq_circuit_to_dem failed (code {rc})")
    try:
        sizes = (c_uint32 * 6)()
        _check(lib.gq_dem_build_sizes(h, ctypes.byref(sizes)), "gq_dem_build_sizes")
        D, O, M, nnz_h, nnz_l, ncomp = (int(x) for x in sizes)
        probs = np.zeros(M, np.float64)
        det_indptr = np.zeros(M + 1, np.uint32); det_idx = np.zeros(nnz_h, np.uint32)
        obs_indptr = np.zeros(M + 1, np.uint32); obs_idx = np.zeros(nnz_l, np.uint32)
        comps = None
        if decompose_errors:
            comps = (np.zeros(M + 1, np.uint32), np.zeros(ncomp, np.uint32), np.zeros(ncomp, np.uint32), np.zeros(ncomp, np.uint64))
        ptrs = [_np_ptr(a) for a in (probs, det_indptr, det_idx, obs_indptr, obs_idx)]
        ptrs += [_np_ptr(a) for a in comps] if comps else [None] * 4
        _check(lib.gq_dem_build_export(h, *ptrs), "gq_dem_build_export")
    finally:
        lib.gq_dem_build_destroy(h)
    return D, O, probs, det_indptr, det_idx, obs_indptr, obs_idx, comps


def _np_ptr(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(c_void_p)


def _torch():
    import torch

    return torch


def _dptr(t):
    return None if t is None else c_void_p(t.data_ptr())


class DemHandle:
    """Device-resident detector error model (``gq_dem``)."""

    def __init__(self, dem, device: int = 0) -> None:
        lib = load_library()
        self.dem = dem
        self.device = int(device)
        self._h = c_void_p()
        probs = np.ascontiguousarray(dem.probs, dtype=np.float64)
        comp = dem.decomposed
        arrs = [
            probs, dem.det_indptr, dem.det_idx, dem.obs_indptr, dem.obs_idx,
            dem.comp_indptr if comp else None, dem.comp_a if comp else None,
            dem.comp_b if comp else None, dem.comp_obs if comp else None,
        ]
        self._keep = arrs
        rc = lib.gq_dem_create(
            dem.num_detectors, dem.num_observables, dem.num_errors,
            *[_np_ptr(a) for a in arrs], self.device, ctypes.byref(self._h),
        )
        _check(rc, "gq_dem_create")
        info = _DemInfo()
        _check(lib.gq_dem_get_info(self._h, ctypes.byref(info)), "gq_dem_get_info")
        self.info = info
        self.D, self.O, self.M = info.num_detectors, info.num_observables, info.num_errors
        self.DW, self.OW, self.tile_shots = info.det_words, info.obs_words, info.tile_shots

    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value:
            load_library().gq_dem_destroy(self._h)
            self._h = c_void_p()

    def __del__(self) -> None:
        try:
            self.close()
        except Exception:   # interpreter shutdown: module globals are already gone
            pass

    # ---- helpers
    def _dev(self):
        return _torch().device("cuda", self.device)

    def _pre(self):
        _torch().cuda.synchronize(self.device)

    def sync(self) -> None:
        _check(load_library().gq_sync(self._h), "gq_sync")

    # ---- K1+K2
    def sample(self, seed: int, shot0: int, nshots: int, return_errors: bool = False):
        """-> (dets_sm int32[nshots, DW], obs_sm int32[nshots, OW][, errs_mm int32[M, SW]])"""
        torch = _torch()
        dets = torch.empty((nshots, self.DW), dtype=torch.int32, device=self._dev())
        obs = torch.empty((nshots, self.OW), dtype=torch.int32, device=self._dev())
        errs = None
        if return_errors:
            errs = torch.zeros((max(self.M, 1), (nshots + 31) // 32), dtype=torch.int32, device=self._dev())
        self._pre()
        rc = load_library().gq_sample(self._h, seed, shot0, nshots, _dptr(dets), _dptr(obs), _dptr(errs))
        _check(rc, "gq_sample")
        self.sync()
        return (dets, obs, errs) if return_errors else (dets, obs)

    def extract(self, errs_mm, nshots: int):
        """Bit-sliced H.e / L.e: errs_mm int32[M, SW] -> (dets_dm int32[D, SW], obs_dm int32[O, SW])"""
        torch = _torch()
        sw = (nshots + 31) // 32
        dets = torch.empty((self.D, sw), dtype=torch.int32, device=self._dev())
        obs = torch.empty((max(self.O, 1), sw), dtype=torch.int32, device=self._dev())
        self._pre()
        rc = load_library().gq_extract(self._h, _dptr(errs_mm), nshots, _dptr(dets), _dptr(obs))
        _check(rc, "gq_extract")
        self.sync()
        return dets, obs

    def dm_to_sm(self, dm, nbits: int, nshots: int):
        torch = _torch()
        sm = torch.zeros((nshots, (nbits + 31) // 32), dtype=torch.int32, device=self._dev())
        self._pre()
        _check(load_library().gq_transpose_dm_to_sm(self._h, _dptr(dm), nbits, nshots, _dptr(sm)), "gq_transpose_dm_to_sm")
        self.sync()
        return sm

    def sm_to_dm(self, sm, nbits: int, nshots: int):
        torch = _torch()
        dm = torch.zeros((nbits, (nshots + 31) // 32), dtype=torch.int32, device=self._dev())
        self._pre()
        _check(load_library().gq_transpose_sm_to_dm(self._h, _dptr(sm), nbits, nshots, _dptr(dm)), "gq_transpose_sm_to_dm")
        self.sync()
        return dm

    def count_errors(self, pred_sm, obs_sm) -> int:
        torch = _torch()
        cnt = torch.zeros(1, dtype=torch.int64, device=self._dev())
        self._pre()
        _check(load_library().gq_count_errors(self._h, _dptr(pred_sm), _dptr(obs_sm), pred_sm.shape[0], _dptr(cnt)), "gq_count_errors")
        self.sync()
        return int(cnt.item())


class DecoderHandle:
    """Device-resident decoder (``gq_dec``): exact matching or min-sum BP."""

    def __init__(self, demh: DemHandle, kind: int = MATCHING, weight_levels: int = 0, fast_capacity: int = 0,
                 with_corrections: bool = False, bp_max_iter: int = 0, bp_alpha: float = 0.0,
                 legacy_tiers: bool = False, bp_osd: bool = False, bp_scratch: int = 0) -> None:
        lib = load_library()
        self.demh = demh
        self.kind = kind
        self._h = c_void_p()
        p = _Params()
        p.weight_levels, p.fast_capacity = int(weight_levels), int(fast_capacity)
        p.with_corrections = int(bool(with_corrections))
        p.bp_max_iter, p.bp_alpha = int(bp_max_iter), float(bp_alpha)
        p.legacy_tiers = int(bool(legacy_tiers))
        p.bp_osd = int(bool(bp_osd))
        p.bp_scratch = int(bp_scratch)      # 0 auto, 1 global-scratch kernel, 2 compressed min-sum kernel
        _check(lib.gq_decoder_create(demh._h, kind, ctypes.byref(p), ctypes.byref(self._h)), "gq_decoder_create")
        gi = _GraphInfo()
        _check(lib.gq_decoder_get_graph_info(self._h, ctypes.byref(gi)), "gq_decoder_get_graph_info")
        self.graph_info = gi
        self.num_edges = gi.num_edges

    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value:
            load_library().gq_decoder_destroy(self._h)
            self._h = c_void_p()

    def __del__(self) -> None:
        try:
            self.close()
        except Exception:   # interpreter shutdown: module globals are already gone
            pass

    def edges(self):
        """-> (u, v [NO_DET = boundary], integer weight, observable mask, merged probability)"""
        n = self.num_edges
        u = np.zeros(n, np.uint32); v = np.zeros(n, np.uint32); w = np.zeros(n, np.uint32)
        om = np.zeros(n, np.uint64); pr = np.zeros(n, np.float64)
        _check(load_library().gq_decoder_get_edges(self._h, _np_ptr(u), _np_ptr(v), _np_ptr(w), _np_ptr(om), _np_ptr(pr)), "gq_decoder_get_edges")
        return u, v, w, om, pr

    def decode(self, dets_sm, want_weights: bool = False, want_flags: bool = False, want_corrections: bool = False):
        torch = _torch()
        n = dets_sm.shape[0]
        dev = dets_sm.device
        pred = torch.zeros((n, self.demh.OW), dtype=torch.int32, device=dev)
        if n == 0:
            want = [want_weights, want_flags, want_corrections]
            extra = [torch.zeros(0, dtype=torch.int32, device=dev), torch.zeros(0, dtype=torch.uint8, device=dev),
                     torch.zeros((0, 1), dtype=torch.int32, device=dev)]
            out = [pred] + [e for e, w in zip(extra, want) if w]
            return out[0] if len(out) == 1 else tuple(out)
        wts = torch.zeros(n, dtype=torch.int32, device=dev) if want_weights else None
        flags = torch.zeros(n, dtype=torch.uint8, device=dev) if want_flags else None
        corr = None
        if want_corrections:
            width = self.num_edges if self.kind == MATCHING else self.demh.M
            corr = torch.zeros((n, (width + 31) // 32), dtype=torch.int32, device=dev)
        self.demh._pre()
        rc = load_library().gq_decode(self._h, _dptr(dets_sm), n, _dptr(pred), _dptr(wts), _dptr(flags), _dptr(corr))
        _check(rc, "gq_decode")
        self.demh.sync()
        out = [pred]
        if want_weights: out.append(wts)
        if want_flags: out.append(flags)
        if want_corrections: out.append(corr)
        return out[0] if len(out) == 1 else tuple(out)

    def decode_dm(self, dets_dm, nshots: int, want_corrections: bool = False):
        """Detector-major in (int32[D, ceil(nshots/32)], what ``DemHandle.extract`` returns), prediction out in the
        same bit-sliced layout (int32[max(O,1), ceil(nshots/32)]); with ``want_corrections`` also the corrections,
        shot-major as ``decode`` returns them."""
        torch = _torch()
        sw = (int(nshots) + 31) // 32
        pred = torch.zeros((max(self.demh.O, 1), sw), dtype=torch.int32, device=dets_dm.device)
        corr = None
        if want_corrections:
            words = (self.graph_info.num_edges + 31) // 32 if self.kind == MATCHING else (self.demh.M + 31) // 32
            corr = torch.zeros((int(nshots), max(1, words)), dtype=torch.int32, device=dets_dm.device)
        if nshots == 0:
            return (pred, corr) if want_corrections else pred
        self.demh._pre()
        _check(load_library().gq_decode_dm(self._h, _dptr(dets_dm), int(nshots), _dptr(pred), _dptr(corr)), "gq_decode_dm")
        self.demh.sync()
        return (pred, corr) if want_corrections else pred

    def decode_b8(self, dets_b8: np.ndarray) -> np.ndarray:
        """Host in, host out, sinter's bit-packed layout (== decode_shots_bit_packed)."""
        dets_b8 = np.ascontiguousarray(dets_b8, dtype=np.uint8)
        n = dets_b8.shape[0]
        db = (self.demh.D + 7) // 8
        if dets_b8.ndim != 2 or dets_b8.shape[1] != db:
            raise ValueError(f"expected uint8[shots, {db}] detection event data, got {dets_b8.shape}")
        ob = max(1, (self.demh.O + 7) // 8)
        pred = np.zeros((n, ob), dtype=np.uint8)
        _check(load_library().gq_decode_b8(self._h, _np_ptr(dets_b8), n, _np_ptr(pred)), "gq_decode_b8")
        return pred

    def collect(self, seed: int, shot0: int, nshots: int, max_errors: int = 0) -> tuple[int, int, int]:
        out = (c_uint64 * 3)()
        rc = load_library().gq_collect(self.demh._h, self._h, seed, shot0, nshots, max_errors, ctypes.byref(out))
        _check(rc, "gq_collect")
        return int(out[0]), int(out[1]), int(out[2])

    def last_kernel_profile(self) -> dict:
        t = (c_double * 4)()
        _check(load_library().gq_last_kernel_profile(self._h, ctypes.byref(t)), "gq_last_kernel_profile")
        return {"ms": t[0], "launches": int(t[1]), "shots": int(t[2]), "K": int(t[3])}

    def last_timing(self) -> dict:
        t = (c_double * 8)()
        _check(load_library().gq_last_timing(self._h, ctypes.byref(t)), "gq_last_timing")
        return {"sample_ms": t[0], "decode_ms": t[1], "count_ms": t[2], "total_ms": t[3],
                "h2d_ms": t[4], "d2h_ms": t[5], "launches": int(t[6])}
