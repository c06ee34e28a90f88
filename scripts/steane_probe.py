"""BASELINE configs[3]: Steane CssCode memory experiment, 1e9 shots per point, low-p tail."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from graphqec_b200 import CssCode, ThresholdLAB, wilson_interval
H = np.array([[1, 1, 1, 1, 0, 0, 0], [0, 1, 1, 0, 1, 1, 0], [0, 0, 1, 1, 0, 1, 1]])
shots = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10**9
th = ThresholdLAB(configurations=[{"Hx": H, "Hz": H}], code=CssCode, error_rates=[1e-4, 3e-4, 1e-3], decoder="pymatching")
t0 = time.time()
th.collect_stats(max_shots=shots, max_errors=10**12, decoder_params={"batch_shots": 1 << 26})
dt = time.time() - t0
for s in th.samples:
    lo, hi = wilson_interval(s.errors, s.shots)
    print(f"{s.json_metadata['name']} p={s.json_metadata['error']:.0e} shots={s.shots:.3e} errors={s.errors} LER={s.errors / s.shots:.3e} [{lo:.3e}, {hi:.3e}] {s.seconds:.2f}s -> {s.shots / s.seconds:.3e} shots/s")
print(f"total {dt:.1f}s")
