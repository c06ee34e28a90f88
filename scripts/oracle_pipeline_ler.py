"""The oracle pipeline (CPU sampler + exact MWPM + count) at the north star, several independent seeds.
python scripts/oracle_pipeline_ler.py shots_per_seed seed [seed ...]"""
import math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("GQEC_PY_DEM", "1")
import bench
from oracle import pyoracle as po
dem = bench.north_star_dem()
orc = po.MwpmOracle(dem)
n = int(float(sys.argv[1]))
for seed in [int(x) for x in sys.argv[2:]]:
    t = time.time()
    shots, errors = po.collect(dem, seed=seed, nshots=n, nthreads=os.cpu_count() or 8, oracle=orc, chunk=1 << 16)
    p = errors / shots
    print("oracle pipeline seed %d: %d / %d = %.4e +- %.2e (%.0f s)" % (seed, errors, shots, p, math.sqrt(p / shots), time.time() - t), flush=True)
