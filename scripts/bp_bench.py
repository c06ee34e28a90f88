"""BASELINE configs[4]: bivariate bicycle [[144,12,12]], 12 rounds -- min-sum BP throughput / convergence on one GPU."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from graphqec_b200 import BivariateBicycleCode, backend as be
from graphqec_b200.lab.threshold.threshold_lab import _auto_dem
p = float(sys.argv[1]) if len(sys.argv) > 1 else 0.001
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 30
osd = (sys.argv[4] == "osd") if len(sys.argv) > 4 else False
mode = int(sys.argv[5]) if len(sys.argv) > 5 else 0      # 0 auto (compressed min-sum kernel), 1 shot-interleaved global-scratch kernel
t0 = time.time()
code = BivariateBicycleCode(depolarize1_rate=p, depolarize2_rate=p)
code.build_memory_circuit(number_of_rounds=12, logic_check="Z")
t1 = time.time()
dem = _auto_dem(code.memory_circuit)
t2 = time.time()
print("circuit %.2fs dem %.2fs" % (t1 - t0, t2 - t1), dem.summary())
demh = be.DemHandle(dem)
bp = be.DecoderHandle(demh, be.BP_MINSUM, bp_max_iter=iters, bp_osd=osd, bp_scratch=mode)
dets, obs = demh.sample(5, 0, n)
import torch
for rep in range(2):
    torch.cuda.synchronize(); t = time.time()
    pred, it, flags = bp.decode(dets, want_weights=True, want_flags=True)
    torch.cuda.synchronize(); dt = time.time() - t
it = it.cpu().numpy(); flags = flags.cpu().numpy()
err = (pred.cpu().numpy()[:, 0] != obs.cpu().numpy()[:, 0])
print(("BP+OSD-0 " if osd else "BP ") + "kernel mode %d " % mode + "p=%g shots=%d: %.1f ms -> %.3e shots/s; valid correction %.4f, mean iters (converged) %.2f, LER all %.4f, LER converged-only %.5f" % (
    p, n, dt * 1e3, n / dt, (flags == 0).mean(), it[it > 0].mean() if (it > 0).any() else 0, err.mean(), err[flags == 0].mean()))
