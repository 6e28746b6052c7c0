#!/usr/bin/env python
"""C4 shape (4,096 runs x 8 agents, 1001x101 fp32 tables in HBM): agent-steps/s of the HBM kernel per gather mode / resident warps,
and of the general kernel, on one GPU.  `python scripts/quick_hbm.py [epochs] [runs] [case ...]`  (case = K=V,K=V)"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch

import bench
from th_rl_b200 import _lib, engine

E = int(sys.argv[1]) if len(sys.argv) > 1 else 100
R = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
cfg = bench.WORKLOADS["c4"]["config"]
hp = bench._c4_hp(R, 8)
cases = [dict(), dict(THRL_HBM_GATHER="bulk"), dict(THRL_HBM_GATHER="bulk", THRL_HBM_LPR="4"), dict(THRL_KERNEL="generic")]
if len(sys.argv) > 3:
    cases = [dict(kv.split("=") for kv in a.split(",") if kv) for a in sys.argv[3:]]
KEYS = ("THRL_HBM_GATHER", "THRL_HBM_NB", "THRL_HBM_LPR", "THRL_HBM_PF", "THRL_HBM_WARPS", "THRL_KERNEL", "THRL_HBM_NOCNT")
for env in cases:
    for k in KEYS:
        os.environ.pop(k, None)
    os.environ.update(env)
    b = engine.RunBatch(cfg, R, seed=0, hp=hp).init_device()
    if env.get("THRL_HBM_NOCNT"):
        b.counter = None
    b.scan(30)  # warm the greedy cache / move down the epsilon schedule a little
    torch.cuda.synchronize()
    t = time.perf_counter()
    b.scan(E)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    print("%-55s %-8s wave %5d  %.3e agent-steps/s  (%.1f ms, %d epochs)" % (env, _lib.last_kernel(), _lib.last_wave_runs(), R * 8 * E * 100 / dt, dt * 1e3, E), flush=True)
    del b
    torch.cuda.empty_cache()
