#!/usr/bin/env python
"""A bench line for BASELINE configs[4]: bivariate bicycle [[144,12,12]], 12 rounds, min-sum BP (30 iterations) -- the same
JSON schema as bench.py's line (which stays on the north-star workload), so the config the metric is NOT quoted on has a
measured line too.

A step = one pass of the hot path (gq_collect: sample -> extract -> BP decode [-> OSD-0] -> count, all on the device) over
`--shots` synthetic shots.  `roofline` is reported for `bp_cms_kernel` against HBM as the contract asks (algorithmic bytes =
the bit-packed syndrome + observable per shot); the kernel keeps a shot's whole state in shared memory and is bound by
instruction issue and shared-memory bandwidth (profiles/r02_bp_cms_kernel_summary.txt), which `secondary_bound` restates.
`cpu_baseline` = oracle/bp_oracle.c (the same min-sum schedule, one thread per shot) on the host cores.

  python scripts/bench_bb144.py [--p 0.003] [--shots 32768] [--steps 3] [--warmup 3] [--osd]
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# profiles/r02_bp_cms_kernel_summary.txt (1184 shots, p = 3e-3, 30 iterations)
NCU_WARP_INSTR_PER_SHOT = 17929008113 / 1184
NCU_ISSUE_ACTIVE_PCT = 70.1
NCU_SHARED_PIPE_PCT = 81.1
NCU_DRAM_BYTES_PER_LAUNCH = 1.9e6


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--p", type=float, default=0.003)
    ap.add_argument("--shots", type=int, default=32768)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--osd", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()

    import numpy as np
    import torch

    import __graft_entry__ as ge
    import bench

    ge.build()
    from graphqec_b200 import BivariateBicycleCode, backend as be
    from graphqec_b200.lab.threshold.threshold_lab import _auto_dem

    code = BivariateBicycleCode(depolarize1_rate=a.p, depolarize2_rate=a.p)
    code.build_memory_circuit(number_of_rounds=12, logic_check="Z")
    dem = _auto_dem(code.memory_circuit)
    assert (dem.num_detectors, dem.num_errors) == (1728, 67752)
    demh = be.DemHandle(dem)
    dec = be.DecoderHandle(demh, be.BP_MINSUM, bp_max_iter=30, bp_osd=a.osd)
    tile = demh.tile_shots
    shots = max(tile, a.shots - a.shots % tile)
    for w in range(max(a.warmup, 3)):
        dec.collect(7, (1000 + w) * shots, shots, 0)
    torch.cuda.synchronize()
    tot = err = disc = 0
    dev_ms = dec_ms = 0.0
    launches = 0
    with bench.ClockSampler(0) as clk:
        t0 = time.perf_counter()
        for k in range(a.steps):
            n, e, d = dec.collect(7, k * shots, shots, 0)
            t = dec.last_timing()
            tot += n; err += e; disc += d
            dev_ms += t["total_ms"]; dec_ms += t["decode_ms"]; launches += t["launches"]
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    value = tot / dt
    alg = (dem.num_detectors + dem.num_observables) / 8.0
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    per_launch_ms = dec_ms / max(1, a.steps)
    achieved = alg * shots / max(per_launch_ms * 1e-3, 1e-12) / 1e9
    cpu = None
    if not a.no_cpu_baseline:
        from oracle import pyoracle as po

        orc = po.BpOracle(dem, max_iter=30, alpha=0.9)
        threads = max(1, os.cpu_count() or 2)
        nb = 16 * threads
        dets, _ = po.sample(dem, 11, nb)
        bits = np.unpackbits(dets, axis=1, bitorder="little")[:, : dem.num_detectors]
        f = (lambda s: orc.decode_osd(bits[s], max_candidates=-2048)[1]) if a.osd else (lambda s: orc.decode(bits[s])[1])
        t0 = time.perf_counter()
        with cf.ThreadPoolExecutor(threads) as ex:      # ctypes releases the GIL inside liboracle.so
            list(ex.map(f, range(nb)))
        cdt = time.perf_counter() - t0
        cpu = {"value": nb / cdt, "unit": "shots/s", "cores": threads, "kind": "port",
               "sample": f"{nb} shots of the same DEM, oracle/bp_oracle.c min-sum BP (30 iterations{', + OSD-0' if a.osd else ''}), {threads} threads, {cdt:.1f} s; "
                         "decode only (the GPU number includes sampling and counting)"}
    line = {
        "metric": "decoded shots/sec", "value": value, "unit": "shots/s", "n_gpus": 1, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1e3 * dt / max(1, a.steps), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BivariateBicycleCode [[144,12,12]] r=12 p=%g Z-memory (BASELINE configs[4])" % a.p,
                   "decoder": "min-sum BP, 30 iterations, alpha 0.9" + (" + OSD-0" if a.osd else ""), "shots_per_step": shots},
        "run": {"logical_errors": err, "discards": disc, "logical_error_rate": err / max(1, tot),
                "device_ms_per_step": dev_ms / max(1, a.steps), "decode_ms_per_step": dec_ms / max(1, a.steps),
                "note": "without OSD an unconverged shot keeps BP's last hard decision (the LER is then that of BP alone)"},
        "gpu_launches": int(launches), "clocks": clk.summary(),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": NCU_DRAM_BYTES_PER_LAUNCH, "kernel": "bp_cms_kernel (K3b)" + (" + osd_rref_kernel" if a.osd else ""),
                     "secondary_bound": {"what": "instruction issue + shared-memory bandwidth: a shot's whole BP state lives in shared memory",
                                         "warp_instr_per_shot": NCU_WARP_INSTR_PER_SHOT, "issue_active_pct": NCU_ISSUE_ACTIVE_PCT,
                                         "shared_pipe_pct": NCU_SHARED_PIPE_PCT, "source": "profiles/r02_bp_cms_kernel_summary.txt"}},
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))


if __name__ == "__main__":
    main()
