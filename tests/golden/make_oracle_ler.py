"""Generate logical-error-rate fixtures with the CPU oracle (sampler + exact MWPM).

Run here (CPU only, minutes); writes tests/golden/oracle_ler.json.  The GPU tests compare their own
counts, at their own shot numbers, against the 95 % Wilson interval of these.
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import graphqec_b200 as gq
from oracle import pyoracle as po

CONFIGS = [
    ("RepetitionCode", {"distance": 3}, 0.1, 200000),
    ("RepetitionCode", {"distance": 5}, 0.1, 200000),
    ("RepetitionCode", {"distance": 7}, 0.05, 200000),
    ("RotatedSurfaceCode", {"distance": 3}, 0.005, 400000),
    ("RotatedSurfaceCode", {"distance": 5}, 0.005, 400000),
    ("RotatedSurfaceCode", {"distance": 5}, 0.008, 200000),
    ("RotatedSurfaceCode", {"distance": 7}, 0.005, 400000),
    ("UnrotatedSurfaceCode", {"distance": 3}, 0.005, 300000),
    ("UnrotatedSurfaceCode", {"distance": 5}, 0.008, 100000),
    ("RotatedSurfaceCode", {"distance": 9}, 0.005, 200000),
    ("RotatedSurfaceCode", {"distance": 11}, 0.005, 600000),
    # (the north-star point at >= 10^7 shots: make_oracle_ler_fine.py -> oracle_ler_fine.json)
    ("RotatedSurfaceCode", {"distance": 11}, 0.008, 300000),     # 129 defects per shot: classes K = 4 .. 6
    # BASELINE configs[3]: Steane code through CssCode, low-p tail (the matching graph ignores the hyperedges, as PyMatching does)
    ("CssCode", {"Hx": [[1, 1, 1, 1, 0, 0, 0], [0, 1, 1, 0, 1, 1, 0], [0, 0, 1, 1, 0, 1, 1]],
                 "Hz": [[1, 1, 1, 1, 0, 0, 0], [0, 1, 1, 0, 1, 1, 0], [0, 0, 1, 1, 0, 1, 1]]}, 0.001, 4000000),
]
nthreads = int(sys.argv[1]) if len(sys.argv) > 1 else 6
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_ler.json")
res = json.load(open(path)) if os.path.exists(path) else []
have = {(r["code"], json.dumps(r["kwargs"]), r["p"], r["shots"]) for r in res}
for cname, kw, p, shots in CONFIGS:
    if (cname, json.dumps(kw), p, shots) in have:
        continue
    import numpy as np
    code = getattr(gq, cname)(**{k: (np.array(v) if isinstance(v, list) else v) for k, v in kw.items()}, depolarize1_rate=p, depolarize2_rate=p)
    code.build_memory_circuit(number_of_rounds=code.distance, logic_check="Z")
    try:
        dem = code.memory_circuit.detector_error_model(decompose_errors=True)
    except ValueError:
        dem = code.memory_circuit.detector_error_model()
    t = time.time()
    n, e = po.collect(dem, seed=987654321, nshots=shots, nthreads=nthreads)
    lo, hi = po.wilson(e, n)
    res.append({"code": cname, "kwargs": kw, "p": p, "rounds": code.distance, "logic_check": "Z", "decoder": "exact MWPM oracle",
                "shots": n, "errors": e, "wilson95": [lo, hi], "seconds": round(time.time() - t, 1)})
    print(res[-1], flush=True)
    json.dump(res, open(path, "w"), indent=1)
