"""One warm-up scan + one measured scan of the C2 shape, for ncu (development helper)."""
import json, sys
import numpy as np, torch
sys.path.insert(0, ".")
from th_rl_b200 import engine
cfg = json.loads(str(np.load("tests/golden/c1_example_2q_seed0.npz")["config"]))
R = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
E = int(sys.argv[2]) if len(sys.argv) > 2 else 50
dtype = torch.float64 if (len(sys.argv) > 3 and sys.argv[3] == "f64") else torch.float32
b = engine.RunBatch(cfg, R, dtype=dtype, seed=0).init_device()
b.scan(E, stats=True)
torch.cuda.synchronize()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record(); b.scan(E, stats=True); t1.record(); torch.cuda.synchronize()
print(json.dumps(dict(R=R, E=E, ms=t0.elapsed_time(t1), agent_steps_per_s=R * 2 * E * 100 / t0.elapsed_time(t1) * 1e3)))
