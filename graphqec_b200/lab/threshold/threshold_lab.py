"""``ThresholdLAB``: sweep (configuration x error rate), decode on the GPU, count logical failures.

Public surface mirrors ``/root/reference/graphqec/lab/threshold/threshold_lab.py:36-267`` --
constructor ``:52-59``, properties ``:81-116`` (``logic_check`` is a *method* there, kept so),
``generate_sinter_tasks`` ``:118-151``, ``collect_stats`` ``:153-197``, ``plot_stats`` ``:199-259``,
``check_is_family`` ``:261-267``.  The reference's ``collect_stats`` is three lines of glue around
``sinter.collect``; here its body is the driver loop over ``libgqec_b200.so``: per task the circuit
becomes a DEM (host), the DEM and its decoder become device-resident handles, and batches of shots
run sample -> extract -> decode -> count entirely on the GPU (``gq_collect``) until ``max_shots`` or
``max_errors`` is reached -- sinter's stop rule.

Multi-GPU: when ``torch.distributed`` is initialised, every batch is cut into one contiguous,
tile-aligned range of shots per rank and the (shots, errors, discards) triple is all-reduced; the
RNG is keyed by global shot index, so totals do not depend on the number of ranks.
"""

from __future__ import annotations

import multiprocessing
import time
import warnings

import numpy as np

from graphqec_b200.codes.base_code import BaseCode
from graphqec_b200.stats import Task, TaskStats

__all__ = ["ThresholdLAB"]

# decoder name -> kernel family.  "pymatching" / "fusion_blossom" are exact MWPM in the reference
# (threshold_lab.py:27-33); "bposd" (stimbposd.SinterDecoder_BPOSD there: product-sum BP + OSD-CS) maps to the
# min-sum BP the BASELINE north star names, followed by order-0 OSD on the shots BP leaves unconverged.
__available_decoders__ = {
    "pymatching": "matching",
    "fusion_blossom": "matching",
    "bposd": "bp_osd",          # min-sum BP, then OSD-0 on the shots BP leaves unconverged
    "bp_minsum": "bp_minsum",   # BP alone (unconverged shots keep BP's last hard decision)
    "mwpf": None,
    "beliefmatching": None,
}

DEFAULT_SEED = 20261019


def _auto_dem(circuit):
    """sinter's rule: decomposed DEM if it exists, else the plain one."""
    try:
        return circuit.detector_error_model(decompose_errors=True)
    except ValueError:
        return circuit.detector_error_model(decompose_errors=False)


def _dist_info():
    try:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            return dist, dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    return None, 0, 1


class ThresholdLAB:
    """Threshold estimation over a family of codes."""

    __slots__ = (
        "_configurations", "_error_rates", "_code", "_collected_stats", "_decoder", "_samples",
        "_code_name", "_logic_check",
    )

    def __init__(self, code: BaseCode, configurations, error_rates, decoder: str = "pymatching",
                 logic_check: str = "Z") -> None:
        self._configurations = configurations
        self._code = code
        self._error_rates = error_rates
        # The reference builds a ValueError for an unknown decoder without raising it
        # (threshold_lab.py:73-76); the name is then simply never stored.  Same here.
        if decoder in __available_decoders__:
            self._decoder = decoder
        self._collected_stats = {}
        self._logic_check = logic_check

    @property
    def configurations(self):
        return self._configurations

    @property
    def error_rates(self):
        return self._error_rates

    @property
    def code(self):
        return self._code

    @property
    def decoder(self) -> str:
        return self._decoder

    @property
    def samples(self):
        return self._samples

    def logic_check(self) -> str:
        return self._logic_check

    # ------------------------------------------------------------------ task generation (A2)
    def generate_sinter_tasks(self, logic_check: str = "Z"):
        """One ``Task`` per (configuration, error rate); rounds = code distance."""
        for configuration in self.configurations:
            for prob_error in self.error_rates:
                code = self.code(
                    **configuration, depolarize1_rate=prob_error, depolarize2_rate=prob_error
                )
                num_rounds = code.distance
                if logic_check not in code.logic_check.keys():
                    raise ValueError(
                        f"Logic check {logic_check} is not supported by {code.name} code."
                    )
                code.build_memory_circuit(number_of_rounds=num_rounds, logic_check=logic_check)
                metadata = {
                    "name": code.name,
                    "error": prob_error,
                    "n": int(code.num_physical_qubits),
                    "k": int(code.num_logical_qubits),
                    "d": int(code.distance),
                    "r": num_rounds,
                }
                yield Task(circuit=code.memory_circuit, json_metadata=metadata)

    # ------------------------------------------------------------------ the hot path (A3)
    def collect_stats(self, num_workers: int | None = None, max_shots: int = 10**4,
                      max_errors: int = 1000, decoder_params: dict | None = None,
                      logic_check: str = "Z", *, save_resume_filepath: str | None = None) -> None:
        """Run every task until ``max_shots`` shots or ``max_errors`` failures; fills ``samples``.

        ``num_workers`` is accepted for signature compatibility (sinter's CPU worker count,
        threshold_lab.py:170-171); the GPU path uses one process per GPU instead.
        ``decoder_params`` may carry backend options: ``seed``, ``device``, ``batch_shots``,
        ``weight_levels``, ``bp_max_iter``, ``bp_alpha``.

        ``save_resume_filepath`` (keyword only; the argument of ``sinter.collect`` the reference never
        wired up, threshold_lab.py:182-189): a CSV in sinter's schema.  Rows already in it count towards
        ``max_shots`` / ``max_errors`` of the task with the same ``strong_id``; what this call adds is
        appended as one row per task (rank 0 writes).  New shots continue the task's counter-based RNG stream
        behind the recorded ones: every row carries the stream position it stopped at (``next_shot`` in sinter's
        ``custom_counts`` column, a whole number of sampler tiles), so a resumed sweep never re-uses a shot.
        """
        from graphqec_b200 import backend as be

        if num_workers is None:
            num_workers = multiprocessing.cpu_count() - 1
        opts = dict(decoder_params or {})
        family = __available_decoders__[self.decoder]
        if family is None:
            raise NotImplementedError(
                f"decoder {self.decoder!r} has no B200 kernel; use 'pymatching' (exact matching), 'bposd' (min-sum BP + OSD-0) or 'bp_minsum'"
            )
        kind = be.MATCHING if family == "matching" else be.BP_MINSUM
        bp_osd = family == "bp_osd"
        seed = int(opts.get("seed", DEFAULT_SEED))
        dist, rank, world = _dist_info()
        device = int(opts.get("device", _default_device(rank)))
        max_shots = int(max_shots)
        max_errors = int(max_errors)

        save_resume_filepath = save_resume_filepath or opts.get("save_resume_filepath")
        prior = {}
        if save_resume_filepath:
            from graphqec_b200.stats import append_stats_to_csv, read_stats_from_csv

            prior = {st.strong_id: st for st in read_stats_from_csv(save_resume_filepath)}

        samples = []
        for task_index, task in enumerate(self.generate_sinter_tasks(logic_check=logic_check)):
            t0 = time.perf_counter()
            dem = _auto_dem(task.circuit)
            shots = errors = discards = 0
            strong_id = task.strong_id(self.decoder)
            before = prior.get(strong_id)
            todo_shots, todo_errors, shot0 = max_shots, max_errors, 0
            if before is not None:
                todo_shots = max(0, max_shots - before.shots)
                if max_errors > 0:
                    todo_errors = max_errors - before.errors
                    if todo_errors <= 0:
                        todo_shots = 0
                # position in the RNG stream: recorded by our own rows; a foreign CSV only has the shot count
                shot0 = int(before.custom_counts.get("next_shot", before.shots))
            next_shot = shot0
            if todo_shots <= 0:
                pass
            elif dem.num_errors == 0:
                # p = 0: nothing can fire, nothing to decode (SURVEY 8a, A6 note on p = 0)
                shots = todo_shots
            else:
                demh = be.DemHandle(dem, device=device)
                dech = be.DecoderHandle(
                    demh, kind,
                    weight_levels=int(opts.get("weight_levels", 0)),
                    bp_max_iter=int(opts.get("bp_max_iter", 0)),
                    bp_alpha=float(opts.get("bp_alpha", 0.0)),
                    bp_osd=bp_osd,
                )
                try:
                    shots, errors, discards = _run_task(
                        dech, demh.tile_shots, seed + 1_000_003 * task_index, todo_shots, todo_errors,
                        int(opts.get("batch_shots", 0)), dist, rank, world, device, shot0=shot0,
                    )
                    tile = demh.tile_shots
                    next_shot = -(-shot0 // tile) * tile + -(-shots // tile) * tile
                finally:
                    dech.close()
                    demh.close()
            new = TaskStats(
                strong_id=strong_id,
                decoder=self.decoder,
                json_metadata=task.json_metadata,
                shots=shots,
                errors=errors,
                discards=discards,
                seconds=time.perf_counter() - t0,
                custom_counts={"next_shot": int(next_shot)} if save_resume_filepath else {},
            )
            if save_resume_filepath and rank == 0 and shots > 0:
                append_stats_to_csv(save_resume_filepath, new)
            if before is not None:
                new.shots += before.shots
                new.errors += before.errors
                new.discards += before.discards
                new.seconds += before.seconds
            samples.append(new)
        self._samples = samples

    # ------------------------------------------------------------------ plotting (A11, host only)
    def plot_stats(self, x_min=None, x_max=None, y_min=None, y_max=None,
                   pseudo_threshold: bool = False, method: str = "per_shot") -> None:
        """Log-log plot of logical error rate against physical error rate (needs matplotlib)."""
        if method not in ("per_shot", "per_round"):
            raise ValueError(
                f"Unknown method: {method}. Method must be one of ['per_shot', 'per_round']."
            )
        try:
            import matplotlib.pyplot as plt
        except ImportError as exc:
            raise RuntimeError("plot_stats needs matplotlib, which is not installed") from exc

        fig, ax = plt.subplots(1, 1)
        for name, xs, ys, lo, hi in self.error_rate_curves(method=method):
            ax.errorbar(xs, ys, yerr=[np.subtract(ys, lo), np.subtract(hi, ys)], marker="o", label=name)
        ax.set_ylabel("Logical Error Rate per Shot" if method == "per_shot" else "Logical Error Rate per Round")
        if pseudo_threshold:
            if not self.check_is_family(self.samples):
                warnings.warn("Inconsistent sample families detected.")
            else:
                k = self.samples[0].json_metadata["k"]
                ax.plot(self.error_rates, k * np.asarray(self.error_rates))
        if x_min is not None and x_max is not None:
            ax.set_xlim(x_min, x_max)
        if y_min is not None and y_max is not None:
            ax.set_ylim(y_min, y_max)
        ax.loglog()
        ax.set_xlabel("Physical Error Rate")
        ax.grid(which="major")
        ax.grid(which="minor")
        ax.legend()
        plt.show()

    def error_rate_curves(self, method: str = "per_shot"):
        """Plot data without matplotlib: ``(name, x, y, y_low, y_high)`` per code, Wilson 95 % bands."""
        from graphqec_b200.stats import wilson_interval

        groups: dict[str, list] = {}
        for st in self.samples:
            groups.setdefault(st.json_metadata["name"], []).append(st)
        out = []
        for name, stats in groups.items():
            stats = sorted(stats, key=lambda s: s.json_metadata["error"])
            xs, ys, lo, hi = [], [], [], []
            for st in stats:
                kept = st.shots - st.discards
                if kept <= 0 or st.errors == 0:
                    continue
                pieces = st.json_metadata["r"] if method == "per_round" else 1
                a, b = wilson_interval(st.errors, kept)
                conv = (lambda v: _shot_to_piece(v, pieces))
                xs.append(st.json_metadata["error"])
                ys.append(conv(st.errors / kept))
                lo.append(conv(a))
                hi.append(conv(b))
            out.append((name, xs, ys, lo, hi))
        return out

    @staticmethod
    def check_is_family(samples) -> bool:
        """True when the samples do *not* all share one name (the reference's sense, ``:261-267``)."""
        first_name = samples[0].json_metadata["name"]
        return any(sample.json_metadata["name"] != first_name for sample in samples)


def _shot_to_piece(rate: float, pieces: int) -> float:
    """sinter.shot_error_rate_to_piece_error_rate: per-shot failure rate -> per-round rate."""
    if pieces == 1:
        return rate
    if rate > 0.5:
        return 1 - _shot_to_piece(1 - rate, pieces)
    return 0.5 - 0.5 * (1 - 2 * rate) ** (1.0 / pieces)


def _default_device(rank: int) -> int:
    import os

    if "LOCAL_RANK" in os.environ:
        return int(os.environ["LOCAL_RANK"])
    try:
        import torch

        n = torch.cuda.device_count()
        return rank % n if n else 0
    except ImportError:
        return 0


def _split_batch(shot0: int, nshots: int, tile: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous tile-aligned share of [shot0, shot0 + nshots) for ``rank`` of ``world``."""
    tiles = (nshots + tile - 1) // tile
    lo_t = tiles * rank // world
    hi_t = tiles * (rank + 1) // world
    lo = min(nshots, lo_t * tile)
    hi = min(nshots, hi_t * tile)
    return shot0 + lo, hi - lo


def _run_task(dech, tile: int, seed: int, max_shots: int, max_errors: int, batch_shots: int,
              dist, rank: int, world: int, device: int, shot0: int = 0) -> tuple[int, int, int]:
    """Batches of shots until sinter's stop rule fires; returns global (shots, errors, discards).

    ``shot0``: shots of this task's RNG stream already consumed by an earlier, recorded run (resume);
    rounded up to a whole tile, the unit the counter-based sampler is keyed by."""
    shots = errors = discards = 0
    base = -(-int(shot0) // tile) * tile
    # ramp the batch so that an early max_errors stop does not overshoot by much
    batch = batch_shots if batch_shots > 0 else 1 << 14
    batch = max(tile, batch - batch % tile)
    cap = batch_shots if batch_shots > 0 else (1 << 24) * world
    while shots < max_shots and (max_errors <= 0 or errors < max_errors):
        n = min(batch, max_shots - shots)
        my0, myn = _split_batch(base + shots, n, tile, rank, world)
        # shot0 handed to the kernel must be tile aligned: ``shots`` only ever grows by whole
        # batches, and batches are multiples of the tile except for the very last one
        got = (0, 0, 0)
        if myn > 0:
            got = dech.collect(seed, my0, myn, 0)
        if dist is not None and world > 1:
            import torch

            dev = torch.device("cuda", device) if dist.get_backend() == "nccl" else torch.device("cpu")
            t = torch.tensor(list(got), dtype=torch.int64, device=dev)
            dist.all_reduce(t)
            got = tuple(int(x) for x in t.tolist())
        shots += got[0]
        errors += got[1]
        discards += got[2]
        if got[0] == 0:
            break
        batch = min(cap, batch * 4)
        batch -= batch % tile
    return shots, errors, discards
