"""Run one BASELINE config through the CUDA path: set-up times, decode throughput, optional oracle parity on a few shots."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import graphqec_b200 as g
from graphqec_b200 import backend as be
from graphqec_b200.lab.threshold.threshold_lab import _auto_dem
name, d, p, n = sys.argv[1], int(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4])
npar = int(sys.argv[5]) if len(sys.argv) > 5 else 0
t0 = time.time()
if name == "toric":
    from graphqec_b200.codes._lattice import toric_checks   # noqa
code = getattr(g, name)(distance=d, depolarize1_rate=p, depolarize2_rate=p)
code.build_memory_circuit(number_of_rounds=d, logic_check="Z")
t1 = time.time(); dem = _auto_dem(code.memory_circuit); t2 = time.time()
s = dem.summary(); print(code.name, "circuit %.2fs dem %.2fs" % (t1 - t0, t2 - t1), {k: s[k] for k in ("D", "M", "nnz_H", "sum_p")}, "decomposed", dem.decomposed, flush=True)
demh = be.DemHandle(dem); t3 = time.time(); dech = be.DecoderHandle(demh, be.MATCHING); t4 = time.time()
print("decoder create %.2fs" % (t4 - t3), flush=True)
dets, obs = demh.sample(5, 0, n)
nd = np.unpackbits(dets.cpu().numpy().view(np.uint8), axis=1).sum(1)
print("defects/shot mean %.1f max %d" % (nd.mean(), nd.max()), flush=True)
for rep in range(2):
    torch.cuda.synchronize(); t = time.time()
    pred, wts, flags = dech.decode(dets, want_weights=True, want_flags=True)
    torch.cuda.synchronize(); dt = time.time() - t
err = (pred.cpu().numpy()[:, 0] != obs.cpu().numpy()[:, 0]).mean()
print("decode %d shots: %.1f ms -> %.3e shots/s, flags!=0: %d, LER %.5f" % (n, dt * 1e3, n / dt, int((flags != 0).sum()), err), flush=True)
torch.cuda.synchronize(); t = time.time(); pred2 = dech.decode(dets); torch.cuda.synchronize(); dt = time.time() - t
print("throughput variant: %.1f ms -> %.3e shots/s, same predictions: %s" % (dt * 1e3, n / dt, bool((pred2 == pred).all())), flush=True)
if npar:
    from oracle import pyoracle as po
    u, v, w, om, pr = dech.edges()
    edges = [(int(a), -1 if int(b) == be.NO_DET else int(b), float(q), int(o)) for a, b, q, o in zip(u, v, pr, om)]
    orc = po.MwpmOracle(edges=edges, num_detectors=dem.num_detectors, num_observables=1, int_weights=w.astype(np.int64))
    b8 = dets[:npar].cpu().numpy().view(np.uint8).reshape(npar, -1)[:, : (dem.num_detectors + 7) // 8]
    t = time.time(); opred, owts = orc.decode_batch(np.ascontiguousarray(b8), nthreads=16, return_weights=True); dt = time.time() - t
    print("oracle %d shots in %.1fs; weights equal: %s" % (npar, dt, bool((wts[:npar].cpu().numpy() == owts).all())), flush=True)
