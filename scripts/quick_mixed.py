"""Device-side timing of games that run on the order-exact MLP kernel (thrl_scan_mixed.cuh): `python scripts/quick_mixed.py
[runs] [epochs] [case]`, case = noisy_aa (C5 agents with demand noise: continuous states), noisy_qr (the shipped example with
noise), cac (two CAC agents), qr_forced (the shipped example, THRL_KERNEL=mixed)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from th_rl_b200 import _lib, engine

def cfg_of(case):
    q = dict(name="QTable", gamma=0.95, actions=21, states=100, alpha=0.1, eps_end=0.001, epsilon=0.5, eps_step=0.9995, action_range=[0.2, 0.4])
    r = dict(name="Reinforce", gamma=0.995, actions=21, states=1, action_range=[0.2, 0.4])
    a = dict(name="ActorCritic", gamma=0.98, actions=21, states=1, action_range=[0.2, 0.4])
    c = dict(name="CAC", gamma=0.98, states=1, action_range=[0.2, 0.4])
    agents, noise = {"noisy_aa": ([a, a], 0.05), "noisy_qr": ([q, r], 0.05), "cac": ([c, c], 0.0), "qr_forced": ([q, r], 0.0)}[case]
    return {"agents": [dict(x) for x in agents],
            "environment": dict(name="NoisyPriceState", noise_prob=noise, a=10, b=1, nplayers=2, max_steps=100),
            "training": dict(print_freq=500, epochs=20)}

R = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
E = int(sys.argv[2]) if len(sys.argv) > 2 else 20
case = sys.argv[3] if len(sys.argv) > 3 else "noisy_aa"
if case == "qr_forced":
    os.environ["THRL_KERNEL"] = "mixed"
b = engine.RunBatch(cfg_of(case), R, seed=0).init_device()
b.scan(10)
torch.cuda.synchronize()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record(); b.scan(E, stats=True); t1.record(); torch.cuda.synchronize()
ms = t0.elapsed_time(t1)
print(json.dumps(dict(case=case, kernel=_lib.last_kernel(), R=R, E=E, ms=ms, agent_steps_per_s=R * 2 * E * 100 / ms * 1e3)))
