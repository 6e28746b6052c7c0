"""ms per launch of the C5 shape over consecutive launches: `python scripts/quick_c5.py [runs] [epochs] [launches] [noise]`."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import bench
from th_rl_b200 import _lib, engine
R = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
E = int(sys.argv[2]) if len(sys.argv) > 2 else 200
L = int(sys.argv[3]) if len(sys.argv) > 3 else 8
noise = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0
b = engine.RunBatch(bench._c5_cfg(E, noise), R, seed=0).init_device()
for i in range(L):
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(); o = b.scan(E, stats=True); t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1)
    st = o.stats.cpu().numpy()
    print("launch %d epochs %d-%d kernel %s: %.1f ms  %.3e agent-steps/s  mean reward %.4f  nan %d" % (
        i, i * E, (i + 1) * E, _lib.last_kernel(), ms, R * 2 * E * 100 / ms * 1e3, st[-1, 0, 0] / 2**32 / R, int(torch.isnan(b.mlp).sum())), flush=True)
