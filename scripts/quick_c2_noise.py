"""The C2 game with the environment's default demand noise (0.05): which kernel plays it and how fast.  `python scripts/quick_c2_noise.py [runs] [epochs]`"""
import os, sys, copy
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import bench
from th_rl_b200 import _lib, engine
R = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
E = int(sys.argv[2]) if len(sys.argv) > 2 else 200
for noise in (0.0, 0.05):
    cfg = copy.deepcopy(bench.WORKLOADS["c2"]["config"])
    cfg["environment"]["noise_prob"] = noise
    b = engine.RunBatch(cfg, R, seed=0).init_device()
    b.scan(E, stats=True); torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(); b.scan(E, stats=True); t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1)
    print("noise %.2f: kernel %s  %.1f ms  %.3e agent-steps/s" % (noise, _lib.last_kernel(), ms, R * 2 * E * 100 / ms * 1e3), flush=True)
