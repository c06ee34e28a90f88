"""Event statistics of the first-tier solver at the north star, from the one-lane host build (scaffolding)."""
import ctypes, os, sys, subprocess
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from oracle import pyoracle as po
import test_oracle as to

so = os.path.join(ROOT, "tests", "hostsim", "libmthostsim.so")
subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-DGQ_FP2_MAXK=1000000", "-o", so, os.path.join(ROOT, "tests", "hostsim", "mt_hostsim.cpp")] + sys.argv[2:])
lib = ctypes.CDLL(so)
lib.hostsim_mt_decode.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_long] + [ctypes.c_void_p] * 5
lib.hostsim_mt_stats.argtypes = [ctypes.c_void_p, ctypes.c_int]
dem = bench.north_star_dem()
orc = po.MwpmOracle(dem)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
dets, _ = po.sample(dem, 3, n)
stats = (ctypes.c_longlong * 16)()
lib.hostsim_mt_stats(stats, 1)
pred, wt, rc = to._tree_decode(lib, orc, dets)
lib.hostsim_mt_stats(stats, 0)
opred, owt = orc.decode_batch(dets, nthreads=8, return_weights=True)
ok = rc == 0
print("rc in {0,4}:", set(rc.tolist()) <= {0, 4}, "decoded", ok.sum(), "weights equal on those", (wt[ok] == owt[ok]).all(), "pred diff", (pred != opred[:, 0]).mean() if opred.ndim > 1 else (pred != opred).mean())
names = ["events(search)", "stale pops", "grow", "shrink", "augment", "4:?", "scans", "expand", "free after start", "shots", "-", "phases with 1 event", "phases with 2 events", "phases with 3 events", "phases with 4+ events", "augments ending at a free vertex"]
shots = max(1, stats[9])
for k, nm in enumerate(names):
    print("  %-20s %.3f per shot" % (nm, stats[k] / shots))
# per class
nd = np.unpackbits(dets, axis=1, bitorder="little").sum(1)
for lo, hi in ((33, 64), (65, 96), (97, 128)):
    sel = (nd >= lo) & (nd <= hi)
    if not sel.any(): continue
    lib.hostsim_mt_stats(stats, 1)
    to._tree_decode(lib, orc, dets[sel])
    lib.hostsim_mt_stats(stats, 0)
    s = max(1, stats[9])
    print(f"class {lo}..{hi}: {sel.sum()} shots, mean n {nd[sel].mean():.1f}: " + ", ".join(f"{nm} {stats[k]/s:.2f}" for k, nm in enumerate(names[:16]) if nm != "-"))
