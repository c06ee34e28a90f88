"""LER of one rotated-surface-code point at many shots: python scripts/ler_point.py d p shots [seed]"""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import graphqec_b200 as gq
from graphqec_b200 import backend as be
d, p, n = int(sys.argv[1]), float(sys.argv[2]), int(float(sys.argv[3]))
seed = int(sys.argv[4]) if len(sys.argv) > 4 else 20261019
code = gq.RotatedSurfaceCode(distance=d, depolarize1_rate=p, depolarize2_rate=p)
code.build_memory_circuit(number_of_rounds=d, logic_check="Z")
dem = code.memory_circuit.detector_error_model(decompose_errors=True)
h = be.DemHandle(dem); dec = be.DecoderHandle(h, be.MATCHING)
shots = errors = 0
step = 1 << 26
k = 0
while shots < n:
    s, e, _ = dec.collect(seed, k * step, min(step, n - shots + (-(n - shots)) % h.tile_shots), 0)
    shots += s; errors += e; k += 1
q = errors / shots
print("rot%d p=%g seed %d: %d / %d = %.5e +- %.1e" % (d, p, seed, errors, shots, q, math.sqrt(q / shots)), flush=True)
