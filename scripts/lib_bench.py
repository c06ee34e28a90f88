import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graphqec_b200 import backend as be
libpath = os.path.abspath(sys.argv[1])
be.library_path = lambda: libpath
import bench
lbp = float(os.environ.get("LB_P", 0.005)); lbd = int(os.environ.get("LB_D", 11))
if lbp == 0.005 and lbd == 11:
    dem = bench.north_star_dem()
else:
    from graphqec_b200 import RotatedSurfaceCode
    code = RotatedSurfaceCode(distance=lbd, depolarize1_rate=lbp, depolarize2_rate=lbp); code.build_memory_circuit(number_of_rounds=lbd, logic_check="Z")
    dem = code.memory_circuit.detector_error_model(decompose_errors=True)
h = be.DemHandle(dem); d = be.DecoderHandle(h, be.MATCHING)
n = int(os.environ.get("LB_SHOTS", 1 << 20))
d.collect(1, 0, n, 0)
ts = []; ss = []
for k in range(3):
    res = d.collect(1, (k + 1) * n, n, 0); ts.append(d.last_timing()["decode_ms"]); ss.append(d.last_timing().get("sample_ms", 0.0))
print(os.path.basename(libpath), res, "decode ms per 2^20:", ["%.1f" % t for t in ts], "-> %.3e shots/s" % (n / (min(ts) * 1e-3)), "sample ms", ["%.2f" % x for x in ss])
