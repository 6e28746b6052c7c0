#!/usr/bin/env python
"""One warm-up launch, then one short launch of the C4 shape for ncu (`-k regex:qtable_scan_hbm -s 1 -c 1`)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch

import bench
from th_rl_b200 import engine

R = int(sys.argv[1]) if len(sys.argv) > 1 else 1036
W = int(sys.argv[2]) if len(sys.argv) > 2 else 30
E = int(sys.argv[3]) if len(sys.argv) > 3 else 3
b = engine.RunBatch(bench.WORKLOADS["c4"]["config"], R, seed=0, hp=bench._c4_hp(R, 8)).init_device()
b.scan(W)
torch.cuda.synchronize()
b.scan(E)
torch.cuda.synchronize()
print("ok", R, W, E)
