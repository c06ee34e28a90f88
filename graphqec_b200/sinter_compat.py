"""sinter-compatible adapters: ``B200Decoder`` / ``B200Sampler``.

The reference forwards ``decoder_params`` to ``sinter.collect(custom_decoders=...)``
(``/root/reference/graphqec/lab/threshold/threshold_lab.py:182-189``); sinter's plug-in contract is
``Decoder.compile_decoder_for_dem(*, dem) -> CompiledDecoder`` with
``CompiledDecoder.decode_shots_bit_packed(*, bit_packed_detection_event_data)``.
These classes implement that contract over ``libgqec_b200.so``.  They subclass sinter's base
classes when sinter is importable and are plain duck-typed classes otherwise.
"""

from __future__ import annotations

import time

import numpy as np

from graphqec_b200 import backend as be
from graphqec_b200.dem import DetectorErrorModel
from graphqec_b200.stats import AnonTaskStats

try:  # pragma: no cover - sinter is absent on the B200 image
    import sinter as _sinter

    _DecoderBase, _CompiledBase = _sinter.Decoder, _sinter.CompiledDecoder
    _SamplerBase, _CompiledSamplerBase = _sinter.Sampler, _sinter.CompiledSampler
except ImportError:
    _DecoderBase = _CompiledBase = _SamplerBase = _CompiledSamplerBase = object

__all__ = ["B200Decoder", "B200CompiledDecoder", "B200Sampler", "B200CompiledSampler"]

_KINDS = {"matching": be.MATCHING, "pymatching": be.MATCHING, "bp_minsum": be.BP_MINSUM, "bposd": be.BP_MINSUM}


def _as_model(dem) -> DetectorErrorModel:
    if isinstance(dem, DetectorErrorModel):
        return dem
    flat = dem.flattened() if hasattr(dem, "flattened") else dem   # stim.DetectorErrorModel
    return DetectorErrorModel.from_text(str(flat))


class B200CompiledDecoder(_CompiledBase):
    def __init__(self, model: DetectorErrorModel, kind: int, device: int, **params) -> None:
        self.model = model
        self.demh = be.DemHandle(model, device=device)
        self.dech = be.DecoderHandle(self.demh, kind, **params)

    def decode_shots_bit_packed(self, *, bit_packed_detection_event_data: np.ndarray) -> np.ndarray:
        """uint8[shots, ceil(D/8)] -> uint8[shots, ceil(O/8)], little-endian bit order (sinter's)."""
        return self.dech.decode_b8(bit_packed_detection_event_data)


class B200Decoder(_DecoderBase):
    """``sinter.Decoder`` backed by the sm_100a kernels; ``kind`` = "matching" or "bp_minsum"."""

    def __init__(self, kind: str = "matching", device: int = 0, **params) -> None:
        if kind == "bposd":
            params.setdefault("bp_osd", True)     # BP, then OSD-0 where BP did not converge
        self.kind, self.device, self.params = _KINDS[kind], device, params

    def compile_decoder_for_dem(self, *, dem) -> B200CompiledDecoder:
        return B200CompiledDecoder(_as_model(dem), self.kind, self.device, **self.params)

    def decode_via_files(self, *, num_shots: int, num_dets: int, num_obs: int, dem_path, dets_b8_in_path,
                         obs_predictions_b8_out_path, tmp_dir=None) -> None:
        """sinter's file form of the plug-in (``sinter.Decoder.decode_via_files``): the DEM as ``.dem`` text,
        detection events as b8 rows (``ceil(num_dets / 8)`` bytes per shot), predictions written as b8 rows
        (``ceil(num_obs / 8)`` bytes per shot)."""
        with open(dem_path) as f:
            model = DetectorErrorModel.from_text(f.read())
        if model.num_detectors > num_dets or model.num_observables > num_obs:
            raise ValueError("the DEM refers to more detectors / observables than the call declares")
        model.num_detectors, model.num_observables = int(num_dets), int(num_obs)
        db, ob = (num_dets + 7) // 8, (num_obs + 7) // 8
        dets = np.fromfile(dets_b8_in_path, dtype=np.uint8)
        if dets.size != num_shots * db:
            raise ValueError(f"{dets_b8_in_path}: expected {num_shots} x {db} bytes, found {dets.size}")
        compiled = B200CompiledDecoder(model, self.kind, self.device, **self.params)
        try:
            pred = compiled.decode_shots_bit_packed(bit_packed_detection_event_data=dets.reshape(num_shots, db))
        finally:
            compiled.dech.close()
            compiled.demh.close()
        out = np.zeros((num_shots, ob), dtype=np.uint8)
        out[:, : pred.shape[1]] = pred[:, :ob]
        out.tofile(obs_predictions_b8_out_path)


class B200CompiledSampler(_CompiledSamplerBase):
    """Fused device path: sample + decode + count per call (sinter >= 1.14 ``CompiledSampler``)."""

    def __init__(self, model: DetectorErrorModel, kind: int, device: int, seed: int, **params) -> None:
        self.demh = be.DemHandle(model, device=device)
        self.dech = be.DecoderHandle(self.demh, kind, **params)
        self.seed, self.next_shot = seed, 0

    def sample(self, suggested_shots: int) -> AnonTaskStats:
        t0 = time.perf_counter()
        tile = self.demh.tile_shots
        n = max(tile, -(-int(suggested_shots) // tile) * tile)
        shots, errors, discards = self.dech.collect(self.seed, self.next_shot, n, 0)
        self.next_shot += n
        return AnonTaskStats(shots=shots, errors=errors, discards=discards, seconds=time.perf_counter() - t0)


class B200Sampler(_SamplerBase):
    def __init__(self, kind: str = "matching", device: int = 0, seed: int = 20261019, **params) -> None:
        if kind == "bposd":
            params.setdefault("bp_osd", True)
        self.kind, self.device, self.seed, self.params = _KINDS[kind], device, seed, params

    def compiled_sampler_for_task(self, task) -> B200CompiledSampler:
        dem = getattr(task, "detector_error_model", None)
        if dem is None:
            circuit = task.circuit
            try:
                dem = circuit.detector_error_model(decompose_errors=True)
            except ValueError:
                dem = circuit.detector_error_model()
        return B200CompiledSampler(_as_model(dem), self.kind, self.device, self.seed, **self.params)
