"""The C2 shape played as one launch or as back-to-back launches over run sub-ranges: `python scripts/quick_c2_split.py`."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import bench
from th_rl_b200 import _lib, engine
R, E = 131072, 200
cfg = bench.WORKLOADS["c2"]["config"]
for piece in (0, 3404, 6808, 34040):
    b = engine.RunBatch(cfg, R, seed=0).init_device()
    def step():
        if piece == 0:
            b.scan(E, stats=True)
        else:
            st = torch.zeros((E, 2, 4), dtype=torch.int64, device=b.device)
            for lo in range(0, R, piece):
                b.scan(E, stats=st, run_range=(lo, min(R, lo + piece)), advance=False)
            b.epoch += E
    step(); step()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(); step(); t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1)
    print("piece %6d: %.1f ms  %.3e agent-steps/s  (%s, wave %d)" % (piece, ms, R * 2 * E * 100 / ms * 1e3, _lib.last_kernel(), _lib.lib().thrl_last_wave_runs()), flush=True)
