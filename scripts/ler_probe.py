"""LER of the north-star point through gq_collect with a given build of the library (A/B of sampler parameters).
python scripts/ler_probe.py build/ab/NAME.so [shots] [seed]"""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graphqec_b200 import backend as be
libpath = os.path.abspath(sys.argv[1])
be.library_path = lambda: libpath
import bench
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1 << 27
seed = int(sys.argv[3]) if len(sys.argv) > 3 else 20261019
dem = bench.north_star_dem()
h = be.DemHandle(dem); d = be.DecoderHandle(h, be.MATCHING)
shots = errors = 0
step = 1 << 24
k = 0
while shots < n:
    s, e, _ = d.collect(seed, k * step, step, 0)
    shots += s; errors += e; k += 1
p = errors / shots
print("%s seed %d: %d / %d = %.4e +- %.2e" % (os.path.basename(libpath), seed, errors, shots, p, math.sqrt(p / shots)), flush=True)
