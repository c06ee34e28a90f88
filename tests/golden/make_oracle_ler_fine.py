"""Offline oracle studies behind the tightened LER acceptance (run HERE, CPU only, ~1 h on 6 threads).

Writes tests/golden/oracle_ler_fine.json with, per point of the rotated d = 11 sweep:

* ``ler``: shots / errors / 95 % Wilson interval of the oracle pipeline (DEM sampler + exact MWPM at the
  product's ``weight_levels = 1000``) at >= 10^7 shots for the north-star point -- the interval the GPU's
  10^8-shot estimate must fall INSIDE (tests/test_gpu_ler.py);
* ``discretisation``: a PAIRED comparison on identical shots of the exact MWPM prediction with integer
  weights at 1000 levels (the product's u16 tables) against 2^20 levels (finer than PyMatching's own
  grid matters for: PyMatching 2 normalises to ~2^24 levels): shots whose prediction differs, and the net
  change in logical errors.  This bounds the bias of the 1000-level grid against an (effectively)
  continuous-weight matcher far more sharply than two independent Wilson intervals could;
* ``ties``: a lower bound on the tie rate -- shots that have two minimum-weight matchings (at 2^14
  levels) with different observable parity, found by decoding with weights 64 w + [edge flips L] and
  64 w - [edge flips L] and comparing predictions at equal base weight.  On such shots "the" MWPM
  prediction is implementation-defined, so exact matchers differ there by chance.
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import graphqec_b200 as gq  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

NTHREADS = int(sys.argv[1]) if len(sys.argv) > 1 else 6
PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_ler_fine.json")
# (distance, p, LER shots, paired-study shots)
POINTS = [(11, 0.005, 12_000_000, 1_000_000), (11, 0.008, 2_000_000, 300_000), (7, 0.005, 4_000_000, 1_000_000)]


def dem_for(d, p):
    code = gq.RotatedSurfaceCode(distance=d, depolarize1_rate=p, depolarize2_rate=p)
    code.build_memory_circuit(number_of_rounds=d, logic_check="Z")
    return code.memory_circuit.detector_error_model(decompose_errors=True)


def oracle_at(dem, edges, levels, bias=0, scale=1):
    w = po.integer_weights(np.array([e[2] for e in edges]), levels=levels) * scale
    if bias:
        w = w + bias * np.array([1 if e[3] else 0 for e in edges], dtype=np.int64)
    return po.MwpmOracle(edges=edges, num_detectors=dem.num_detectors, num_observables=dem.num_observables, int_weights=w), w


def main():
    res = json.load(open(PATH)) if os.path.exists(PATH) else {}
    for d, p, n_ler, n_pair in POINTS:
        key = f"rot{d}_p{p}"
        ent = res.setdefault(key, {"code": "RotatedSurfaceCode", "distance": d, "p": p, "rounds": d, "logic_check": "Z"})
        dem = dem_for(d, p)
        edges = po.merged_edges(dem)
        if "discretisation" not in ent:
            t = time.time()
            o1, _ = oracle_at(dem, edges, 1000)
            o2, _ = oracle_at(dem, edges, 1 << 20)
            diff = e1 = e2 = done = 0
            k = 0
            while done < n_pair:
                n = min(50_000, n_pair - done)
                dets, obs = po.sample(dem, 555_000 + k, n)
                a = o1.decode_batch(dets, nthreads=NTHREADS)
                b = o2.decode_batch(dets, nthreads=NTHREADS)
                diff += int((a != b).sum())
                e1 += int((a != obs[:, 0]).sum())
                e2 += int((b != obs[:, 0]).sum())
                done += n
                k += 1
            ent["discretisation"] = {"shots": done, "levels_a": 1000, "levels_b": 1 << 20, "predictions_differ": diff,
                                     "errors_a": e1, "errors_b": e2, "seconds": round(time.time() - t, 1)}
            print(key, ent["discretisation"], flush=True)
            json.dump(res, open(PATH, "w"), indent=1)
        if "ties" not in ent:
            t = time.time()
            lv = 1 << 14
            op, wp = oracle_at(dem, edges, lv, bias=+1, scale=64)
            om, wm = oracle_at(dem, edges, lv, bias=-1, scale=64)
            n = min(n_pair, 300_000)
            ties = done = 0
            k = 0
            while done < n:
                m = min(50_000, n - done)
                dets, obs = po.sample(dem, 777_000 + k, m)
                a, wa = op.decode_batch(dets, nthreads=NTHREADS, return_weights=True)
                b, wb = om.decode_batch(dets, nthreads=NTHREADS, return_weights=True)
                # base weights: wa = 64 W + (#L-edges used) and wb = 64 W' - (#L-edges used), both counts < 64
                same_base = (wa // 64) == -((-wb) // 64)
                ties += int(((a != b) & same_base).sum())
                done += m
                k += 1
            ent["ties"] = {"shots": done, "levels": lv, "shots_with_parity_tie": ties, "seconds": round(time.time() - t, 1)}
            print(key, ent["ties"], flush=True)
            json.dump(res, open(PATH, "w"), indent=1)
        if "ler" not in ent or ent["ler"]["shots"] < n_ler:
            t = time.time()
            orc = po.MwpmOracle(dem)
            n, e = po.collect(dem, seed=424242, nshots=n_ler, nthreads=NTHREADS, oracle=orc, chunk=50_000)
            lo, hi = po.wilson(e, n)
            ent["ler"] = {"shots": n, "errors": e, "wilson95": [lo, hi], "weight_levels": 1000, "seconds": round(time.time() - t, 1)}
            print(key, ent["ler"], flush=True)
            json.dump(res, open(PATH, "w"), indent=1)


if __name__ == "__main__":
    main()
