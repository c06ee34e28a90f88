"""Paired check at the north star: the same shots decoded by the GPU and by the exact-MWPM oracle (graph built
independently from the DEM), for shots drawn by the GPU sampler and by the oracle's sampler.  Separates
decoder differences (ties) from sampler differences.  python scripts/gpu_vs_oracle_paired.py [shots]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from graphqec_b200 import backend as be
from oracle import pyoracle as po
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1 << 20
dem = bench.north_star_dem()
demh = be.DemHandle(dem); dech = be.DecoderHandle(demh, be.MATCHING)
orc = po.MwpmOracle(dem)
nt = os.cpu_count() or 8
db = (dem.num_detectors + 7) // 8
def b8(t, nbits):
    a = t.cpu().numpy().view(np.uint8).reshape(t.shape[0], -1)
    return np.ascontiguousarray(a[:, : max(1, (nbits + 7) // 8)])
# (a) GPU-sampled shots
dets, obs = demh.sample(4242, 0, n)
pred = dech.decode(dets)
d8, o8, p8 = b8(dets, dem.num_detectors), b8(obs, 1)[:, 0] & 1, b8(pred, 1)[:, 0] & 1
t = time.time(); op = orc.decode_batch(d8, nthreads=nt); dt = time.time() - t
print("GPU-sampled %d shots: GPU errors %d, oracle errors %d, predictions differ on %d shots; mean defects %.2f (oracle %.0f shots/s on %d threads)" % (
    n, int((p8 != o8).sum()), int((op != o8).sum()), int((p8 != op).sum()), np.unpackbits(d8, axis=1).sum() / n, n / dt, nt), flush=True)
# (b) oracle-sampled shots
tot = 0; eg = eo = diff = 0; defs = 0
k = 0
while tot < n:
    m = min(1 << 18, n - tot)
    od, oo = po.sample(dem, 90000 + k, m)
    rows = np.zeros((m, demh.DW * 4), dtype=np.uint8); rows[:, :db] = od
    tdev = torch.from_numpy(rows.view(np.int32)).cuda()
    pg = b8(dech.decode(tdev), 1)[:, 0] & 1
    oq = orc.decode_batch(od, nthreads=nt)
    eg += int((pg != (oo[:, 0] & 1)).sum()); eo += int((oq != (oo[:, 0] & 1)).sum()); diff += int((pg != oq).sum())
    defs += int(np.unpackbits(od, axis=1).sum())
    tot += m; k += 1
print("oracle-sampled %d shots: GPU errors %d, oracle errors %d, predictions differ on %d shots; mean defects %.2f" % (tot, eg, eo, diff, defs / tot), flush=True)
