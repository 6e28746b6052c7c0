#!/usr/bin/env python
"""Reference-ensemble fixture: learning curves of many seeded runs of the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY (needs /root/reference; its output tests/golden/ensemble/*.npz is committed and travels).

The free-running (Philox) mode of the device path cannot be bit-compared with the reference's MT19937 / torch streams
(SURVEY D7, H7), so its parity is distributional: tests/test_gpu_ensemble.py trains a few hundred device runs of the same
config and compares, window by window, the distribution of the per-run mean total reward with this fixture.

Per case: RUNS seeded runs of th_rl/trainer.py:29-110 `train_one` (seed s -> random.seed(s), numpy.random.seed(s),
torch.manual_seed(s)) x EPOCHS epochs; stored: the per-run, per-window (WINDOW epochs) mean of log.csv's rewards and
actions columns, float64 [RUNS, EPOCHS // WINDOW, n].

Usage:  PYTHONDONTWRITEBYTECODE=1 python oracle/make_ensemble.py [--procs 8] [--only NAME]
"""
import argparse
import contextlib
import io
import json
import multiprocessing as mp
import os
import random
import sys
import tempfile

import numpy

REFERENCE = os.environ.get("THRL_REFERENCE", "/root/reference")
RUNS, EPOCHS, WINDOW = 32, 2000, 50


def _q(**kw):
    d = dict(name="QTable", gamma=0.95, actions=21, states=100, alpha=0.1, eps_end=0.001, epsilon=0.5, eps_step=0.9995,
             action_range=[0.2, 0.4])
    d.update(kw)
    return d


_ENV = dict(name="NoisyPriceState", noise_prob=0, a=10, b=1, nplayers=2, max_steps=100)
CASES = {
    # example_config.json hyper-parameters, agent-0 block duplicated (BASELINE C1/C2)
    "ensemble_2q": dict(agents=[_q(), _q()], environment=dict(_ENV), training=dict(epochs=EPOCHS, print_freq=100000)),
    # the shipped example_config.json pairing (th_rl/some_path/configs/example_config.json:2-27)
    "ensemble_qr": dict(agents=[_q(), dict(name="Reinforce", gamma=0.995, actions=21, states=1, action_range=[0.2, 0.4])],
                        environment=dict(_ENV), training=dict(epochs=EPOCHS, print_freq=100000)),
}


def one_run(job):
    name, seed = job
    sys.path.insert(0, REFERENCE)
    sys.dont_write_bytecode = True
    import torch
    import th_rl.trainer as rtrainer
    torch.set_num_threads(1)
    random.seed(seed)
    numpy.random.seed(seed)
    torch.manual_seed(seed)
    cfg = CASES[name]
    with tempfile.TemporaryDirectory() as tmp:
        cpath = os.path.join(tmp, "cfg.json")
        with open(cpath, "w") as f:
            json.dump(cfg, f)
        out = os.path.join(tmp, "run")
        with contextlib.redirect_stdout(io.StringIO()):
            rtrainer.train_one(out, cpath)  # the unmodified reference loop
        with open(os.path.join(out, "log.csv")) as f:
            f.readline(), f.readline()
            log = numpy.loadtxt(f, delimiter=",", ndmin=2)
    n = len(cfg["agents"])
    w = log.reshape(EPOCHS // WINDOW, WINDOW, 2 * n).mean(axis=1)
    return name, seed, w[:, :n], w[:, n:]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "ensemble"))
    ap.add_argument("--procs", type=int, default=os.cpu_count())
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    names = [k for k in CASES if not args.only or k == args.only]
    jobs = [(k, 1000 + s) for k in names for s in range(RUNS)]
    with mp.get_context("spawn").Pool(args.procs) as pool:
        res = pool.map(one_run, jobs, chunksize=1)
    for k in names:
        rows = sorted((r for r in res if r[0] == k), key=lambda r: r[1])
        path = os.path.join(args.out, k + ".npz")
        numpy.savez_compressed(path, config=numpy.array(json.dumps(CASES[k])), seeds=numpy.array([r[1] for r in rows]),
                               window=numpy.int64(WINDOW), rewards=numpy.stack([r[2] for r in rows]),
                               actions=numpy.stack([r[3] for r in rows]))
        tot = numpy.stack([r[2] for r in rows]).sum(-1)
        print("%-14s runs=%d  last-window total reward: median %.3f  IQR %.3f-%.3f  %5.1f KB" % (
            k, len(rows), numpy.median(tot[:, -1]), *numpy.percentile(tot[:, -1], [25, 75]), os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
