"""Device-side timing of the shipped example_config pairing (QTable + Reinforce) batched over R runs."""
import json, sys
import torch
sys.path.insert(0, ".")
from th_rl_b200 import engine
cfg = {"agents": [dict(name="QTable", gamma=0.95, actions=21, states=100, alpha=0.1, eps_end=0.001, epsilon=0.5, eps_step=0.9995, action_range=[0.2, 0.4]),
                  dict(name="Reinforce", gamma=0.995, actions=21, states=1, action_range=[0.2, 0.4])],
       "environment": dict(name="NoisyPriceState", noise_prob=0, a=10, b=1, nplayers=2, max_steps=100),
       "training": dict(print_freq=500, epochs=20)}
R = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
E = int(sys.argv[2]) if len(sys.argv) > 2 else 20
b = engine.RunBatch(cfg, R, seed=0).init_device()
b.scan(10)
torch.cuda.synchronize()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record(); b.scan(E, stats=True); t1.record(); torch.cuda.synchronize()
ms = t0.elapsed_time(t1)
print(json.dumps(dict(shape="example_config (QTable+Reinforce)", R=R, E=E, ms=ms, agent_steps_per_s=R * 2 * E * 100 / ms * 1e3)))
