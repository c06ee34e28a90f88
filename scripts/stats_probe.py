import sys, os, ctypes, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graphqec_b200 import backend as be
libpath = os.path.abspath(sys.argv[1])
be.library_path = lambda: libpath
import bench
dem = bench.north_star_dem()
h = be.DemHandle(dem); d = be.DecoderHandle(h, be.MATCHING)
lib = be.load_library()
lib.gq_debug_stats.argtypes = [ctypes.c_void_p, ctypes.c_int]
out = (ctypes.c_ulonglong * 16)()
lib.gq_debug_stats(out, 1)
n = 1 << 18
res = d.collect(1, 0, n, 0)
lib.gq_debug_stats(out, 1)
names = ["events(search)", "stale pops", "grow", "shrink", "augment", "expand", "scans", "shots", "free after greedy", "eager recomputes"]
print(libpath, res, d.last_timing()["decode_ms"])
shots = out[7]
for k, nm in enumerate(names): print("  %-20s %.2f per shot" % (nm, out[k] / max(1, shots)))
