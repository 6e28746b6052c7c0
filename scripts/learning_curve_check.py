"""Distributional check of the free-running (Philox) mode against the reference's recorded runs.

The reference ships two runs of example_config.json (QTable + Reinforce, 20,000 epochs); their mean total reward per step over
the last 1,000 epochs is 22.35 and 21.51 (th_rl/some_path/runs/example_config/{0,"1 "}/log.csv; Nash 22.22, cartel 25,
th_rl/utils.py:91-92).  This trains the same config for the same number of epochs on R independent runs on the GPU and prints
the same statistic per run (quantiles), plus the 2 x QTable twin."""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from th_rl_b200 import trainer

def cfg(second):
    q = dict(name="QTable", gamma=0.95, actions=21, states=100, alpha=0.1, eps_end=0.001, epsilon=0.5, eps_step=0.9995, action_range=[0.2, 0.4])
    r = dict(name="Reinforce", gamma=0.995, actions=21, states=1, action_range=[0.2, 0.4])
    return {"agents": [q, dict(q) if second == "QTable" else r],
            "environment": dict(name="NoisyPriceState", noise_prob=0, a=10, b=1, nplayers=2, max_steps=100),
            "training": dict(print_freq=500, epochs=20000)}

R = int(sys.argv[1]) if len(sys.argv) > 1 else 64
out = {}
for second in ("Reinforce", "QTable"):
    t = time.time()
    res = trainer.train_many(cfg(second), R, seed=0, log_runs=R, chunk_epochs=2000)
    torch.cuda.synchronize()
    tot = res.rewards_log[:, -1000:, :].sum(2).mean(1)   # per run: mean total reward per step over the last 1000 epochs
    out["QTable+" + second] = dict(runs=R, seconds=round(time.time() - t, 1), mean=float(tot.mean()), sd=float(tot.std()),
                                   quantiles={q: float(np.quantile(tot, q)) for q in (0.05, 0.25, 0.5, 0.75, 0.95)},
                                   first_500_epochs_mean=float(res.rewards_log[:, :500, :].sum(2).mean()))
print(json.dumps(out, indent=1))
