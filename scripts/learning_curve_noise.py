"""Free-running 20,000-epoch learning curves of the shipped example pairing (QTable + Reinforce) with the environment's
default demand noise (noise_prob = 0.05), on the interval-table kernel (default dispatch) and on the order-exact kernel
(THRL_KERNEL=mixed): same Philox streams, so the two must tell the same story run by run.  Prints, per kernel, the per-run mean
total reward per step over the last 1,000 epochs (quantiles), and how far the runs of the two kernels are apart."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from th_rl_b200 import trainer, _lib

def cfg(noise):
    q = dict(name="QTable", gamma=0.95, actions=21, states=100, alpha=0.1, eps_end=0.001, epsilon=0.5, eps_step=0.9995, action_range=[0.2, 0.4])
    r = dict(name="Reinforce", gamma=0.995, actions=21, states=1, action_range=[0.2, 0.4])
    return {"agents": [q, r], "environment": dict(name="NoisyPriceState", noise_prob=noise, a=10, b=1, nplayers=2, max_steps=100),
            "training": dict(print_freq=500, epochs=20000)}

R = int(sys.argv[1]) if len(sys.argv) > 1 else 64
E = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
c = cfg(0.05)
c["training"]["epochs"] = E
out, logs = {}, {}
for kern in ("pwc", "mixed"):
    os.environ.pop("THRL_KERNEL", None)
    if kern == "mixed":
        os.environ["THRL_KERNEL"] = "mixed"
    t = time.time()
    res = trainer.train_many(c, R, seed=0, log_runs=R, chunk_epochs=2000)
    torch.cuda.synchronize()
    assert _lib.last_kernel() == kern, _lib.last_kernel()
    logs[kern] = res.rewards_log
    tot = res.rewards_log[:, -1000:, :].sum(2).mean(1)
    out[kern] = dict(runs=R, epochs=E, seconds=round(time.time() - t, 1), mean=float(tot.mean()), sd=float(tot.std()),
                     quantiles={q: float(np.quantile(tot, q)) for q in (0.05, 0.25, 0.5, 0.75, 0.95)})
a, m = logs["pwc"], logs["mixed"]
same = (a == m).reshape(R, -1).all(axis=1)
first = [int(np.argmax((a[r] != m[r]).any(axis=1))) if not same[r] else E for r in range(R)]
out["pwc_vs_mixed"] = dict(runs_with_bit_identical_reward_logs=int(same.sum()), first_differing_epoch_quantiles={q: float(np.quantile(first, q)) for q in (0.0, 0.25, 0.5, 0.75)},
                           max_abs_difference_of_final_statistic=float(np.abs(a[:, -1000:, :].sum(2).mean(1) - m[:, -1000:, :].sum(2).mean(1)).max()))
print(json.dumps(out, indent=1))
