"""Quick device-side timing of the scan on the C2 shape (development helper, not the bench)."""
import json, sys, time
import torch
sys.path.insert(0, ".")
from th_rl_b200 import engine
import numpy as np

cfg = json.loads(str(np.load("tests/golden/c1_example_2q_seed0.npz")["config"]))
R = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
E = int(sys.argv[2]) if len(sys.argv) > 2 else 50
import os
for dtype in (torch.float32, torch.float64):
    b = engine.RunBatch(cfg, R, dtype=dtype, seed=0).init_device()
    b.scan(2)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(); b.scan(E, stats=True); t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1)
    steps = R * 2 * E * 100
    print(json.dumps(dict(kernel=os.environ.get('THRL_KERNEL','auto'), gl=os.environ.get('THRL_LPC_GL',''), dtype=str(dtype), R=R, E=E, ms=ms, agent_steps_per_s=steps / ms * 1e3)))
