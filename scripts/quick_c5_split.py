"""The C5 shape played as one launch, or as back-to-back launches over run sub-ranges (single rounds of the persistent grid):
`python scripts/quick_c5_split.py [noise]`."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import bench
from th_rl_b200 import _lib, engine
noise = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0
R, E = 16384, 200
for piece in (0, 2368, 2072, 2048, 1776, 1184):
    b = engine.RunBatch(bench._c5_cfg(E, noise), R, seed=0).init_device()
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
    print("piece %5d: %.1f ms  %.3e agent-steps/s  (%s)" % (piece, ms, R * 2 * E * 100 / ms * 1e3, _lib.last_kernel()), flush=True)
