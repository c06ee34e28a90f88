"""GPU run against the statistics the reference itself recorded (tests/golden/notebook_ler_readoffs.json).

For every plot point read off the reference's notebooks and every TaskStats row of its toric notebook, run the same
(code, p, logic_check) through this library at a shot count that makes our own sampling error negligible, and print
ours / theirs / z-score (their binomial error from ``approx_errors`` plus 3 % read-off error).  Writes a JSON log
(default gpurun_out/reference_stats.json) -- the evidence behind tests/test_gpu_reference_stats.py and DESIGN.md section 2.

  python scripts/reference_stats.py [--plots] [--toric] [--max-L 12] [--out PATH]
"""
import argparse
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

import graphqec_b200 as gq  # noqa: E402
from graphqec_b200 import backend as be  # noqa: E402
from make_reference_circuits import materialise  # noqa: E402

SEED = 20261019


def run_point(code_name, spec, p, logic_check, shots, max_errors=0):
    code = getattr(gq, code_name)(**materialise(spec), depolarize1_rate=p, depolarize2_rate=p)
    code.build_memory_circuit(number_of_rounds=code.distance, logic_check=logic_check)
    try:
        dem = code.memory_circuit.detector_error_model(decompose_errors=True)
    except ValueError:
        dem = code.memory_circuit.detector_error_model()
    demh = be.DemHandle(dem, device=0)
    dech = be.DecoderHandle(demh, be.MATCHING)
    n = -(-int(shots) // demh.tile_shots) * demh.tile_shots
    t0 = time.perf_counter()
    got = dech.collect(SEED, 0, n, max_errors)
    dt = time.perf_counter() - t0
    dech.close(); demh.close()
    return got[0], got[1], dt, dem.num_detectors


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--plots", action="store_true")
    ap.add_argument("--toric", action="store_true")
    ap.add_argument("--max-L", type=int, default=12)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "reference_stats.json"))
    args = ap.parse_args()
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "notebook_ler_readoffs.json")))
    out = {"plots": [], "toric": []}
    if args.plots:
        for pt in gold["points"]:
            ler = pt["ler"]
            shots = min(1 << 23, max(1 << 18, int(40000 / ler)))     # ~40 k failures: our own error is ~0.5 %
            n, e, dt, D = run_point(pt["code"], pt["spec"], pt["p"], pt["logic_check"], shots)
            ours = e / n
            rel = math.sqrt(1.0 / max(pt["approx_errors"], 1.0) + 0.03 ** 2 + 1.0 / max(e, 1))
            z = math.log(max(ours, 1e-12) / ler) / rel
            row = dict(pt, shots=n, errors=e, ours=ours, z=z, seconds=dt)
            out["plots"].append(row)
            print("%-22s %-34s %s p=%.5f  ref %.3e (~%4d err)  ours %.3e (%7d err / %8d)  ratio %.3f  z %+5.2f" % (
                pt["code"], json.dumps(pt["spec"]), pt["logic_check"], pt["p"], ler, pt["approx_errors"], ours, e, n, ours / ler, z), flush=True)
            json.dump(out, open(args.out, "w"), indent=1)
    if args.toric:
        for row in gold["toric_taskstats"]:
            if row["L"] > args.max_L:
                continue
            spec = {"toric": row["L"], "name": "Toric"}
            ref = row["errors"] / row["shots"]
            shots = row["shots"] if row["L"] <= 8 else min(row["shots"], 1 << 18 if row["L"] >= 16 else 1 << 20)
            n, e, dt, D = run_point("CssCode", spec, row["p"], "Z", shots)
            ours = e / n
            rel = math.sqrt(1.0 / max(row["errors"], 1) + 1.0 / max(e, 1))
            z = math.log(max(ours, 1e-12) / max(ref, 1e-12)) / rel
            out["toric"].append(dict(row, our_shots=n, our_errors=e, ours=ours, z=z, our_seconds=dt, detectors=D))
            print("toric L=%2d p=%.5f  ref %7d / %7d = %.4e   ours %7d / %7d = %.4e  ratio %.3f  z %+6.2f   %.1f s (ref %.0f core-s)" % (
                row["L"], row["p"], row["errors"], row["shots"], ref, e, n, ours, ours / max(ref, 1e-12), z, dt, row["seconds"]), flush=True)
            json.dump(out, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
