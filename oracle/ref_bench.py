"""Times the UNMODIFIED reference (oracle/_ref, staged by oracle/stage_reference.py) on host cores.

MEASUREMENT INFRASTRUCTURE ONLY: used by bench.py's cpu_baseline and --impl reference legs (SURVEY 8(d) "CPU reference
timing": P processes, each one th_rl/trainer.py:29-110 `train_one` on the workload's config with `epochs` reduced, one
thread each, stdout suppressed, perf_counter around train_one only -- imports and process start-up are excluded).
"""
import contextlib
import io
import json
import multiprocessing as mp
import os
import sys
import tempfile
import time

from . import stage_reference


def _worker(job):
    cfg, seed = job
    os.environ["OMP_NUM_THREADS"] = "1"
    os.environ["MKL_NUM_THREADS"] = "1"
    sys.dont_write_bytecode = True
    sys.path.insert(0, stage_reference.TARGET)
    import random
    import numpy
    import torch
    torch.set_num_threads(1)
    from th_rl.trainer import train_one  # the reference's own entry point
    random.seed(seed), numpy.random.seed(seed), torch.manual_seed(seed)
    with tempfile.TemporaryDirectory() as tmp:
        cpath = os.path.join(tmp, "cfg.json")
        with open(cpath, "w") as f:
            json.dump(cfg, f)
        with contextlib.redirect_stdout(io.StringIO()):
            t = time.perf_counter()
            train_one(os.path.join(tmp, "run"), cpath)
            dt = time.perf_counter() - t
    return dt


def available():
    return stage_reference.staged()


class ReferencePool:
    """`procs` worker processes that have imported the staged reference; every sample() runs one train_one per worker."""

    def __init__(self, procs):
        self.procs = int(procs)
        self.pool = mp.get_context("spawn").Pool(self.procs)
        self.pool.map(_noop, range(self.procs))  # start the workers (imports excluded from the timing, as SURVEY 8(d) says)
        self.calls = 0

    def sample(self, config, epochs):
        """`procs` concurrent train_one calls of `config` with training.epochs = epochs.  Returns (agent-steps/s aggregate,
        seconds of the slowest process, agent-steps per process)."""
        cfg = json.loads(json.dumps(config))
        cfg["training"] = dict(cfg.get("training", {}), epochs=int(epochs), print_freq=10 ** 9)
        per_proc = len(cfg["agents"]) * int(epochs) * int(cfg["environment"]["max_steps"])
        self.calls += 1
        dts = self.pool.map(_worker, [(cfg, 1000 * self.calls + k) for k in range(self.procs)], chunksize=1)
        return per_proc * self.procs / max(dts), max(dts), per_proc

    def close(self):
        self.pool.close()
        self.pool.join()


def time_reference(config, epochs, procs):
    pool = ReferencePool(procs)
    try:
        return pool.sample(config, epochs)
    finally:
        pool.close()


def _noop(_):
    sys.path.insert(0, stage_reference.TARGET)
    import torch  # noqa: F401
    import th_rl.trainer  # noqa: F401
    return 0
