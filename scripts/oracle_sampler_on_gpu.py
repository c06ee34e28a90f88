"""LER of the north-star point with shots drawn by the ORACLE's sampler (oracle/sampler_oracle.c, double precision,
per-mechanism geometric skipping) and decoded on the GPU (the two decoders agree on all but ~5 of 10^6 shots:
scripts/gpu_vs_oracle_paired.py) -- the oracle pipeline's error rate at a shot count the CPU decoder cannot reach,
to compare with the GPU sampler's.  python scripts/oracle_sampler_on_gpu.py [shots] [seed0]"""
import math, os, sys, time
from concurrent.futures import ProcessPoolExecutor
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CHUNK = 1 << 18
_dem = None


def _sample(seed):
    global _dem
    import bench
    from oracle import pyoracle as po
    if _dem is None:
        os.environ["GQEC_PY_DEM"] = "0"
        _dem = bench.north_star_dem()
    d, o = po.sample(_dem, seed, CHUNK)
    return d, o[:, 0] & 1


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1 << 26
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 31337
    import bench
    from graphqec_b200 import backend as be
    dem = bench.north_star_dem()
    demh = be.DemHandle(dem); dech = be.DecoderHandle(demh, be.MATCHING)
    shots = errors = defects = 0
    t0 = time.time()
    with ProcessPoolExecutor(max_workers=max(1, (os.cpu_count() or 4) - 2)) as ex:
        for d8, obs in ex.map(_sample, [seed0 + 1000003 * k for k in range((n + CHUNK - 1) // CHUNK)]):
            pred = dech.decode_b8(d8)[:, 0] & 1
            errors += int((pred != obs).sum()); shots += len(obs)
    p = errors / shots
    print("oracle sampler + GPU decoder, seed0 %d: %d / %d = %.4e +- %.2e  (%.0f s)" % (seed0, errors, shots, p, math.sqrt(p / shots), time.time() - t0), flush=True)


if __name__ == "__main__":
    main()
