"""Device-side timing of the C4 shape (8 agents, 1001x101 tables in HBM, per-run hyper-parameter sweep)."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, ".")
from th_rl_b200 import engine

def c4_config(states=1000, actions=101, n=8, T=100):
    a = dict(name="QTable", gamma=0.95, actions=actions, states=states, alpha=0.1, eps_end=0.001, epsilon=0.5,
             eps_step=0.9995, action_range=[0.05, 0.15])
    return {"agents": [dict(a) for _ in range(n)],
            "environment": dict(name="NoisyPriceState", noise_prob=0, a=10, b=1, nplayers=n, max_steps=T),
            "training": dict(print_freq=500, epochs=200)}

def sweep_hp(R, n):
    grid = [(al, g, st) for al in (.05, .1, .2, .5) for st in (.999, .9995, .9999, .99995) for g in (.35, .8, .95, .99)]
    hp = np.empty((R, n, 4))
    for r in range(R):
        al, g, st = grid[r % 64]
        hp[r, :, 0], hp[r, :, 1], hp[r, :, 2], hp[r, :, 3] = al, g, 0.001, st
    return hp

if __name__ == "__main__":
    R = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    E = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    cfg = c4_config()
    b = engine.RunBatch(cfg, R, dtype=torch.float32, seed=0, hp=sweep_hp(R, 8)).init_device()
    b.scan(2, stats=True)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(); b.scan(E, stats=True); t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1)
    steps = R * 8 * E * 100
    print(json.dumps(dict(shape="c4", R=R, E=E, ms=ms, agent_steps_per_s=steps / ms * 1e3,
                          hbm_GBps_algorithmic=steps * 824 / ms * 1e3 / 1e9)))
