"""BASELINE configs[1] through the reference-facing API: RotatedSurfaceCode d in {5,7,9,11}, d rounds, p in [1e-3, 1e-2],
N shots per point (default 1e8), on the GPU(s) of this process group (python or torchrun).  Writes sinter-schema CSV."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from graphqec_b200 import RotatedSurfaceCode, ThresholdLAB, wilson_interval

shots = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10**8
out = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/sweep_configs1.csv"
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
if world > 1:
    import torch, torch.distributed as dist
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
if rank == 0 and os.path.exists(out):
    os.remove(out)
th = ThresholdLAB(configurations=[{"distance": d} for d in (5, 7, 9, 11)], code=RotatedSurfaceCode,
                  error_rates=np.linspace(1e-3, 1e-2, 10), decoder="pymatching")
t0 = time.time()
th.collect_stats(max_shots=shots, max_errors=10**12, save_resume_filepath=out, decoder_params={"batch_shots": 1 << 24} if world == 1 else None)
dt = time.time() - t0
if rank == 0:
    tot = sum(s.shots for s in th.samples)
    print(f"{len(th.samples)} points, {tot:.3e} shots in {dt:.1f} s on {world} GPU(s) = {tot / dt:.3e} shots/s (set-up included)")
    for s in th.samples:
        lo, hi = wilson_interval(s.errors, s.shots)
        print(f"d={s.json_metadata['d']:2d} p={s.json_metadata['error']:.4f} shots={s.shots} errors={s.errors} LER={s.errors / s.shots:.3e} [{lo:.3e}, {hi:.3e}] {s.seconds:.2f}s")
if world > 1:
    dist.destroy_process_group()
