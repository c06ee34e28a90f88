#!/usr/bin/env python
"""bench.py -- decoded shots/s of the threshold-estimation hot path at the north-star workload.

Workload (BASELINE.json metric, configs[1]): RotatedSurfaceCode d=11, 11 rounds, p=0.005, Z memory;
DEM built by this repo's own builder and asserted against SURVEY.md section 8d's sizes.

A "step" = one pass of the hot path (sample -> extract -> decode -> count, all on the device) over one
batch of `--shots` synthetic shots per GPU.  `value` = whole-job shots / wall time of the K timed
steps (barrier + synchronize on both sides, max over ranks).  `e2e` = the same shots/s measured
through the host-buffer entry point (`gq_decode_b8` == sinter's decode_shots_bit_packed): pinned host
syndromes in, host predictions out, copies inside the timed region, failure count on the host.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--shots S] [--impl reference]
  (N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...)

`--impl reference` times the CPU restatement of the reference's path (oracle/: DEM sampler + exact
MWPM, all host threads) on a bounded sample; the real reference (sinter + Stim + PyMatching wheels)
cannot be installed in this image ("Oracle").
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 20261019
WORKLOAD = "RotatedSurfaceCode d=11 r=11 p=0.005 Z-memory (BASELINE configs[1], north-star point)"
ALG_BYTES_PER_SHOT = (1320 + 1) / 8.0   # bit-packed syndrome + observable per shot
# from the committed ncu --set full capture of the dominant kernel (profiles/, this round); None = not captured
# profiles/r02_tree_solve_kernel3_summary.txt: 465.7 MB read + 7.6 MB written / 766 327 shots (the hand-off
# records are ~6 B per defect + 16 B of that), 3.216e9 warp-instructions, 75.0 % of issue slots active
NCU_DRAM_BYTES_PER_SHOT = (465.733376e6 + 7.633920e6) / 766327
NCU_WARP_INSTR_PER_SHOT = 3215855058 / 766327
NCU_ISSUE_ACTIVE_PCT = 75.0
# both arms print this object verbatim (the driver compares it); run-dependent numbers live under "run"
CONFIG = {"workload": WORKLOAD, "code": "RotatedSurfaceCode", "distance": 11, "rounds": 11, "p": 0.005, "logic_check": "Z",
          "decoder": "exact minimum-weight perfect matching"}
# the reference's own recorded throughput for instances of similar size (sinter + PyMatching, unknown CPU, summed
# worker seconds): /root/reference/notebooks/toric_code.ipynb:184,191
NOTEBOOK_NOTE = ("reference notebook, sinter+PyMatching per core: toric L=8 (D=1024, p=0.006) 18.3 k shots/s, "
                 "toric L=12 (D=3456, p=0.006) 4.3 k shots/s (notebooks/toric_code.ipynb:184,191)")


_JSON_FD = None


def emit_json(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def north_star_dem():
    from graphqec_b200 import RotatedSurfaceCode

    code = RotatedSurfaceCode(distance=11, depolarize1_rate=0.005, depolarize2_rate=0.005)
    code.build_memory_circuit(number_of_rounds=11, logic_check="Z")
    # SURVEY 8d sizes are those of the symptom-merged (= undecomposed) model ...
    s = code.memory_circuit.detector_error_model().summary()
    assert (s["D"], s["O"], s["M"], s["nnz_H"], s["nnz_L"]) == (1320, 1, 24483, 79956, 835), s
    assert abs(s["sum_p"] - 36.94) < 0.005 and abs(s["min_p"] - 3.341e-4) < 1e-7 and abs(s["max_p"] - 2.2227e-2) < 1e-6
    # ... the model sinter asks Stim for is the decomposed one, where every distinct decomposition of a symptom is an
    # error of its own (graphqec_b200/dem.py): same symptoms, same total rates, 27 680 lines
    dem = code.memory_circuit.detector_error_model(decompose_errors=True)
    s = dem.summary()
    assert (s["D"], s["O"], s["M"], s["nnz_H"], s["nnz_L"]) == (1320, 1, 27680, 92594, 871), s
    assert abs(s["sum_p"] - 36.942) < 0.005
    return dem


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region: NVML (pynvml, ~1 ms per sample, every
    10 ms) when importable, else the nvidia-smi query of the profiling recipe (one sample per ~0.2 s)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None
        self._nvml = None
        try:
            import pynvml

            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else index
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self._nvml = pynvml
        except Exception:
            self._nvml = None

    def _sample_nvml(self):
        n = self._nvml
        sm = n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self._h, n.NVML_CLOCK_SM)
        r = n.nvmlDeviceGetCurrentClocksEventReasons(self._h) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") else n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
        flags = []
        for name, bit in (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4)):
            flags.append("Active" if r & bit else "Not Active")
        return [str(sm), str(mx), ""] + flags

    def _run(self):
        while not self._stop.is_set():
            try:
                if self._nvml is not None:
                    self.rows.append(self._sample_nvml())
                else:
                    out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                         capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.01 if self._nvml is not None else 0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows), "source": "nvml" if self._nvml is not None else "nvidia-smi"}


def cpu_baseline(dem, seconds_target: float = 15.0, threads: int | None = None):
    """Oracle port (sample + exact MWPM + count) on the host cores, bounded sample."""
    from oracle import pyoracle as po

    threads = threads or max(1, (os.cpu_count() or 2))
    orc = po.MwpmOracle(dem)
    po.set_jump_start(True)      # the faster schedule of the port (same optimum): the baseline is not handicapped
    t0 = time.perf_counter()
    n, e = po.collect(dem, seed=1, nshots=2000, nthreads=threads, oracle=orc, chunk=2000)
    rate = n / (time.perf_counter() - t0)
    shots = int(max(2000, min(1000000, rate * seconds_target)))
    t0 = time.perf_counter()
    n, e = po.collect(dem, seed=2, nshots=shots, nthreads=threads, oracle=orc, chunk=20000)
    dt = time.perf_counter() - t0
    po.set_jump_start(False)
    return {"value": n / dt, "unit": "shots/s", "cores": threads, "kind": "port",
            "sample": f"{n} shots of the same DEM: oracle DEM sampler + exact blossom MWPM + count, {threads} threads, {dt:.1f} s; {e} logical errors; "
                      f"{n / dt / threads:.0f} shots/s per core -- a dense primal-dual blossom with a greedy dual-feasible start, not PyMatching's sparse blossom; {NOTEBOOK_NOTE}"}, (n, e, dt)


def run_reference(args):
    """The reference's path on the host cores.  Its wheels (stim / sinter / pymatching) cannot be installed in this
    image, so this is the oracle port; the process never loads the product's library: the DEM comes from the
    pure-Python builder (graphqec_b200/dem.py) and everything timed is oracle/liboracle.so."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    os.environ["GQEC_PY_DEM"] = "1"
    dem = north_star_dem()
    from oracle import pyoracle as po

    threads = max(1, (os.cpu_count() or 2))
    orc = po.MwpmOracle(dem)
    po.set_jump_start(True)      # the faster schedule of the port (same optimum weight, tests/test_oracle.py)
    # bounded sample per step so that K + W steps end within a few minutes
    probe_n = 2000
    t0 = time.perf_counter()
    po.collect(dem, seed=3, nshots=probe_n, nthreads=threads, oracle=orc, chunk=probe_n)
    rate = probe_n / (time.perf_counter() - t0)
    budget = 150.0 / max(1, args.steps + args.warmup)
    per_step = int(max(2000, min(1000000, rate * budget)))
    for w in range(args.warmup):
        po.collect(dem, seed=10 + w, nshots=per_step, nthreads=threads, oracle=orc)
    t0 = time.perf_counter()
    shots = errors = 0
    for k in range(args.steps):
        n, e = po.collect(dem, seed=100 + k, nshots=per_step, nthreads=threads, oracle=orc)
        shots += n
        errors += e
    dt = time.perf_counter() - t0
    val = shots / dt
    line = {
        "impl": "reference", "metric": "decoded shots/sec", "value": val, "unit": "shots/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u16", "data": "synthetic",
        "config": dict(CONFIG),
        "run": {"shots_per_step": per_step, "logical_errors": errors, "logical_error_rate": errors / max(1, shots)},
        "cpu_baseline": {"value": val, "unit": "shots/s", "cores": threads, "kind": "port",
                         "sample": f"{per_step} shots/step x {args.steps} steps, oracle port (DEM sampler + exact MWPM + count), {threads} threads, "
                                   f"{val / threads:.0f} shots/s per core (dense primal-dual blossom with a greedy dual-feasible start); the reference's wheels (stim/sinter/pymatching) are not installable here; {NOTEBOOK_NOTE}"},
        "e2e": {"value": val, "unit": "shots/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_json(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--shots", type=int, default=1 << 22, help="shots per step per GPU")
    ap.add_argument("--e2e-shots", type=int, default=0, help="shots per e2e step per GPU (default: --shots)")
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    # stdout carries exactly one JSON line: everything libraries print there (NCCL's version banner, ...) is
    # sent to stderr, and the line itself is written to the saved descriptor
    sys.stdout.flush()
    global _JSON_FD
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"   # NCCL prints its version line to stdout; stdout carries the one JSON line
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    from graphqec_b200 import backend as be

    dem = north_star_dem()
    demh = be.DemHandle(dem, device=local)
    dech = be.DecoderHandle(demh, be.MATCHING)
    tile = demh.tile_shots
    shots = max(tile, args.shots - args.shots % tile)
    dev = torch.device("cuda", local)
    counts = torch.zeros(3, dtype=torch.int64, device=dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step(k: int):
        # rank r owns the shot range [(k*world + r) * shots, +shots): disjoint, counter-based RNG
        got = dech.collect(SEED, (k * world + rank) * shots, shots, 0)
        t = dech.last_timing()
        t["kernel"] = dech.last_kernel_profile()
        return got, t

    for w in range(max(args.warmup, 3) if args.warmup else 0):
        step(10_000 + w)
    barrier()
    tot_shots = tot_err = 0
    dev_ms = dec_ms = samp_ms = 0.0
    launches = 0
    k_ms, k_launches, k_shots, k_K = 0.0, 0, 0, 0
    with ClockSampler(local) as clk:
        t0 = time.perf_counter()
        for k in range(args.steps):
            got, t = step(k)
            tot_shots += got[0]
            tot_err += got[1]
            dev_ms += t["total_ms"]; dec_ms += t["decode_ms"]; samp_ms += t["sample_ms"]
            launches += t["launches"]
            kp = t["kernel"]
            k_ms += kp["ms"]; k_launches += kp["launches"]; k_shots += kp["shots"]; k_K = kp["K"]
        if world > 1:
            # the path's only exchange: one all-reduce of (shots, errors, discards) over NVLink, inside the timed region
            counts.copy_(torch.tensor([tot_shots, tot_err, 0], dtype=torch.int64))
            dist.all_reduce(counts)
            tot_shots, tot_err = (int(x) for x in counts.tolist()[:2])
        barrier()
        dt = time.perf_counter() - t0
    tmax = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dt = float(tmax.item())
    value = tot_shots / dt

    # ---- e2e: host buffers through gq_decode_b8 (the sinter plug-in seam), copies inside the timing
    e2e_shots = args.e2e_shots or args.shots
    ne = max(tile, e2e_shots - e2e_shots % tile)
    dets_dev, obs_dev = demh.sample(SEED + 1, 0, ne)
    db = (dem.num_detectors + 7) // 8
    host_dets = torch.empty((ne, db), dtype=torch.uint8, pin_memory=True)
    host_dets.copy_(dets_dev.view(torch.uint8).reshape(ne, -1)[:, :db])
    host_obs = obs_dev.view(torch.uint8).reshape(ne, -1)[:, :1].cpu().numpy()
    del dets_dev, obs_dev
    hd = host_dets.numpy()
    dech.decode_b8(hd[: tile * 16])
    if args.warmup:
        dech.decode_b8(hd)   # one untimed step of the timed size: the call sizes its device buffers on first use
    barrier()
    e2e_steps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    e2e_err = 0
    e2e_t = {}
    for _ in range(e2e_steps):
        pred = dech.decode_b8(hd)
        e2e_t = dech.last_timing()
        e2e_err = int(np.count_nonzero(pred[:, 0] != host_obs[:, 0])) if pred.shape[1] == 1 else int((pred != host_obs).any(axis=1).sum())
    barrier()
    e2e_dt = time.perf_counter() - t0
    te = torch.tensor([e2e_dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * ne * e2e_steps / float(te.item())

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        # dominant kernel: the blossom-phase kernel of the class most shots fall in (K3a, tree_solve_kernel<K>),
        # timed inside gq_collect with CUDA events on the decoder's stream.  Algorithmic bytes = the
        # bit-packed syndrome + observable of the shots THAT kernel decoded.
        per_launch_ms = k_ms / max(1, k_launches)
        shots_per_launch = k_shots / max(1, k_launches)
        achieved = ALG_BYTES_PER_SHOT * shots_per_launch / max(per_launch_ms * 1e-3, 1e-12) / 1e9
        samp_gbs = ALG_BYTES_PER_SHOT * shots / max(samp_ms / max(1, args.steps) * 1e-3, 1e-12) / 1e9
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                # the fraction the metric text names: sampling / extraction (sample_kernel, K1 + K2 fused with the
                # classification), bit-packed syndrome + observable written once
                "sampler": {"kernel": "sample_kernel (K1+K2)", "achieved": samp_gbs, "frac": samp_gbs / peak, "unit": "GB/s",
                            "shots_per_s": shots / max(samp_ms / max(1, args.steps) * 1e-3, 1e-12),
                            "bound_note": "Philox + log1pf + shared-memory atomicXor issue, not HBM (profiles/r02_sample_kernel_summary.txt)"},
                "whole_step": {"achieved": ALG_BYTES_PER_SHOT * shots / max(dev_ms / max(1, args.steps) * 1e-3, 1e-12) / 1e9,
                               "frac": ALG_BYTES_PER_SHOT * shots / max(dev_ms / max(1, args.steps) * 1e-3, 1e-12) / 1e9 / peak},
                "traffic": NCU_DRAM_BYTES_PER_SHOT * shots_per_launch if NCU_DRAM_BYTES_PER_SHOT else None,
                "kernel": "tree_solve_kernel<%d> (K3a, blossom phases of the shots with %d..%d defects)" % (k_K, 32 * (k_K - 1) + 1, 32 * k_K),
                "peak_source": "MEASURED_PEAKS.json (hbm_gbs, burst)" if peaks else "fallback",
                "kernel_ms_per_launch": per_launch_ms, "shots_per_launch": shots_per_launch,
                "kernel_share_of_step": k_ms / max(dev_ms, 1e-9),
                "kernel_shots_per_s": k_shots / max(k_ms * 1e-3, 1e-12),
                "secondary_bound": {"what": "warp-instruction issue (the bound that binds; DESIGN.md section 4)",
                                    "warp_instr_per_shot": NCU_WARP_INSTR_PER_SHOT, "issue_active_pct": NCU_ISSUE_ACTIVE_PCT,
                                    "source": "profiles/ (ncu --set full of the same kernel)"},
                "note": "exact matching is bound by integer issue / L2 latency, not HBM (DESIGN.md); the HBM fraction is reported as the contract asks"}
        cpu = None
        if not args.no_cpu_baseline:
            cpu, _ = cpu_baseline(dem)
        lo, hi = 0.0, 1.0
        from graphqec_b200 import wilson_interval

        lo, hi = wilson_interval(tot_err, tot_shots)
        line = {
            "metric": "decoded shots/sec", "value": value, "unit": "shots/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(1, args.steps), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u16", "data": "synthetic",
            "config": dict(CONFIG),
            "run": {"shots_per_step_per_gpu": shots, "kernels": "warp-per-shot blossom, per-K start / solve kernels, dynamic work distribution",
                    "l2": "inputs larger than L2: each step streams %d MB of syndromes per GPU" % (shots * 168 // 2**20),
                    "logical_errors": tot_err, "logical_error_rate": tot_err / max(1, tot_shots), "ler_wilson95": [lo, hi],
                    "device_ms_per_step": dev_ms / max(1, args.steps), "sample_ms_per_step": samp_ms / max(1, args.steps),
                    "decode_ms_per_step": dec_ms / max(1, args.steps)},
            "e2e": {"value": e2e_value, "unit": "shots/s", "h2d_bytes_per_step": ne * db, "d2h_bytes_per_step": ne,
                    "api": "gq_decode_b8 (decode_shots_bit_packed layout), pinned host syndromes, host failure count",
                    "shots_per_step": ne, "logical_errors": e2e_err,
                    "h2d_ms": e2e_t.get("h2d_ms"), "decode_ms": e2e_t.get("decode_ms"), "d2h_ms": e2e_t.get("d2h_ms")},
            "gpu_launches": int(launches),
            "clocks": clk.summary(),
            "roofline": roof,
            "cpu_baseline": cpu,
        }
        emit_json(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
