import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graphqec_b200 import backend as be
libpath = os.path.abspath(sys.argv[1])
be.library_path = lambda: libpath
import bench
dem = bench.north_star_dem()
h = be.DemHandle(dem); d = be.DecoderHandle(h, be.MATCHING)
n = int(os.environ.get("LB_SHOTS", 1 << 20))
d.collect(1, 0, n, 0)
ts = []; ss = []
for k in range(3):
    res = d.collect(1, (k + 1) * n, n, 0); ts.append(d.last_timing()["decode_ms"]); ss.append(d.last_timing().get("sample_ms", 0.0))
print(os.path.basename(libpath), res, "decode ms per 2^20:", ["%.1f" % t for t in ts], "-> %.3e shots/s" % (n / (min(ts) * 1e-3)), "sample ms", ["%.2f" % x for x in ss])
