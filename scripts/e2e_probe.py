"""e2e (host buffers through gq_decode_b8) for several chunk schedules: GQ_B8_FIRST_LOG2 / GQ_B8_MAX_LOG2 are read once per
process, so every schedule runs in its own subprocess.  python scripts/e2e_probe.py [shots]"""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, ROOT)
    import numpy as np, torch
    import bench
    from graphqec_b200 import backend as be
    n = int(sys.argv[2])
    dem = bench.north_star_dem()
    demh = be.DemHandle(dem); dech = be.DecoderHandle(demh, be.MATCHING)
    dets, obs = demh.sample(5, 0, n)
    db = (dem.num_detectors + 7) // 8
    host = torch.empty((n, db), dtype=torch.uint8, pin_memory=True)
    host.copy_(dets.view(torch.uint8).reshape(n, -1)[:, :db])
    hd = host.numpy()
    dech.decode_b8(hd)
    best = None
    for _ in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        pred = dech.decode_b8(hd)
        dt = time.perf_counter() - t0
        t = dech.last_timing()
        if best is None or dt < best[0]:
            best = (dt, t)
    dt, t = best
    print("first 2^%s max 2^%s: wall %.2f ms = %.3e shots/s | exposed h2d %.2f decode %.2f d2h %.2f launches %d" % (
        os.environ.get("GQ_B8_FIRST_LOG2"), os.environ.get("GQ_B8_MAX_LOG2"), dt * 1e3, n / dt, t["h2d_ms"], t["decode_ms"], t["d2h_ms"], t["launches"]), flush=True)
else:
    n = sys.argv[1] if len(sys.argv) > 1 else str(1 << 22)
    for f, m in ((16, 19), (17, 19), (17, 20), (18, 20), (17, 21), (18, 21), (18, 22), (22, 22)):
        env = dict(os.environ, GQ_B8_FIRST_LOG2=str(f), GQ_B8_MAX_LOG2=str(m))
        subprocess.run([sys.executable, os.path.abspath(__file__), "--child", n], env=env)
