"""trainer — drop-in for th_rl/trainer.py on the B200 path.

    create_game(configpath)                         th_rl/trainer.py:13-26   same signature, same return triple
    train_one(exp_path, configpath, ...)            th_rl/trainer.py:29-110  same signature, same files on disk
    train_many(config, runs, ...)                   th_rl/main.py:19-21      the reference's `for i in range(runs)` loop,
                                                                             played as ONE batched device scan

`train_one` keeps the reference's random streams (`rng="reference"`, the default): tables and the initial price are drawn
from numpy's global generator by the same constructor calls in the same order, and the exploration / noise draws are taken
from python's `random` / `numpy.random` exactly where th_rl/agents.py:81-82 and th_rl/environments.py:28-29 would take them,
then replayed by the kernel (THRL_RNG_REPLAY_DRAWS) on float64 tables.  A script that seeds `random` and `numpy.random`
therefore gets bit-identical `<i>.npy`, `<i>_counter.npy` and `log.csv` from either implementation
(tests/test_gpu_dropin.py).  `train_many` uses the device's counter-based Philox streams instead.
"""
import json
import os
import random
import time

import numpy

from . import abi
from .agents import AGENTS
from .environments import ENVIRONMENTS


def _load_config(config_or_path):
    if isinstance(config_or_path, (str, os.PathLike)):
        with open(config_or_path) as f:
            return json.load(f)
    return config_or_path


def create_game(configpath):
    """th_rl/trainer.py:13-26.  Class lookup is a registry instead of eval(); unknown names raise NameError like eval would."""
    config = _load_config(configpath)
    agents = []
    for agent in config["agents"]:
        if agent["name"] not in AGENTS:
            raise NameError("name %r is not defined (the reference's agents are QTable, Reinforce, ActorCritic and CAC)" % agent["name"])
        agents.append(AGENTS[agent["name"]](**agent))
    assert len(agents) == config["environment"]["nplayers"], "Bad config. Check number of agents."
    if config["environment"]["name"] not in ENVIRONMENTS:
        raise NameError("name %r is not defined" % config["environment"]["name"])
    environment = ENVIRONMENTS[config["environment"]["name"]](**config["environment"])
    return config, agents, environment


def reference_streams(agents, environment, epochs):
    """Draw, on the host, exactly the random numbers the reference's loop would consume for `epochs` episodes, in its order:
    per step, per agent: u = random.uniform(0,1) then random.choice(action_space) iff u < epsilon (agents.py:81-82, epsilon
    frozen within an episode, decayed after it: agents.py:78); then the environment's numpy.random.uniform(0,1) and, iff it is
    below noise_prob, numpy.random.uniform(0.7a, a) (environments.py:28-29).  Returns (u, ra, new_a) for REPLAY_DRAWS."""
    n, T = len(agents), environment.max_steps
    u = numpy.full((epochs, T, n), numpy.nan, numpy.float64)
    ra = numpy.full((epochs, T, n), -1, numpy.int32)
    new_a = numpy.empty((epochs, T), numpy.float64)
    qt = [hasattr(a, "epsilon") for a in agents]  # MLP agents draw from torch's generator inside the step, not here:
    eps = [a.epsilon if q else 0.0 for a, q in zip(agents, qt)]  # their columns stay (nan, -1) = "sample on the device"
    spaces = [a.action_space if q else None for a, q in zip(agents, qt)]
    a_hi, noise = environment.a, environment.noise_prob
    uniform, choice, npuniform = random.uniform, random.choice, numpy.random.uniform
    for e in range(epochs):
        ue, re_, ne = u[e], ra[e], new_a[e]
        for t in range(T):
            for i in range(n):
                if not qt[i]:
                    continue
                x = uniform(0, 1)
                ue[t, i] = x
                if x < eps[i]:
                    re_[t, i] = choice(spaces[i])
            if npuniform(0, 1) < noise:
                ne[t] = npuniform(a_hi * 0.7, a_hi)
            else:
                ne[t] = a_hi
        eps = [a.eps_end + (x - a.eps_end) * a.eps_step if q else 0.0 for a, x, q in zip(agents, eps, qt)]
    return u, ra, new_a


def _print_progress(config, agents_cfg, rewards_log, actions_log, eps_now, e, print_freq, dt, print_eps):
    # trainer.py:73-98, same format strings
    rew = numpy.mean(rewards_log[e - print_freq + 1: e + 1, :], axis=0)
    act = numpy.mean(actions_log[e - print_freq + 1: e + 1, :], axis=0)
    names = ",".join([a["name"] for a in agents_cfg])
    if print_eps:
        print("eps:{} | time:{:2.2f} | episode:{:3d} | reward:{} | agents:{} | actions:{}".format(
            numpy.round(numpy.array(eps_now) * 1000) / 1000, dt, e, numpy.round(100 * rew) / 100, names,
            numpy.round(100 * act) / 100))
    else:
        print("time:{:2.2f} | episode:{:3d} | reward:{} | agents:{} | actions:{}".format(
            dt, e, numpy.round(100 * rew) / 100, names, numpy.round(100 * act) / 100))


def save_run(exp_path, config, agents, rewards_log, actions_log):
    """trainer.py:101-110: <i>.npy, <i>_counter.npy, config.json (indent 3), log.csv (two header rows, index=None)."""
    import pandas
    if not os.path.exists(exp_path):
        os.mkdir(os.path.join(exp_path))
    for i, a in enumerate(agents):
        a.save(os.path.join(exp_path, str(i)))
    with open(os.path.join(exp_path, "config.json"), "w") as f:
        json.dump(config, f, indent=3)
    rpd = pandas.DataFrame(data=rewards_log, columns=numpy.arange(len(agents)))
    apd = pandas.DataFrame(data=actions_log, columns=numpy.arange(len(agents)))
    log = pandas.concat([rpd, apd], axis=1, keys=["rewards", "actions"])
    log.to_csv(os.path.join(exp_path, "log.csv"), index=None)


def train_one(exp_path, configpath, loadonly=False, print_eps=False, rng="reference", device="cuda:0", chunk_epochs=None):
    """th_rl/trainer.py:29-110 for one run, played on the device.  `loadonly` is accepted and unused, as in the reference."""
    import torch
    from . import engine

    if not os.path.exists(exp_path):
        os.mkdir(os.path.join(exp_path))
    config, agents, environment = create_game(configpath)

    epochs = config.get("training", {}).get("epochs", 0)
    max_steps = config.get("environment", {}).get("max_steps", 0)
    print_freq = config.get("training", {}).get("print_freq", 500)
    n = len(agents)
    rewards_log = numpy.zeros((epochs, n))
    actions_log = numpy.zeros((epochs, n))

    t = time.time()
    state = environment.reset()  # trainer.py:45 — the second uniform draw, like the reference
    is_q = [isinstance(a, AGENTS["QTable"]) for a in agents]
    # Philox key of this run.  In reference mode the QTable / environment draws are replayed from the host generators, but an
    # MLP agent's action samples are drawn on the device (the reference takes them from torch's generator inside every step,
    # agents.py:160-163): their key comes from torch's generator, so a script that seeds torch stays reproducible and
    # successive runs get independent sampling noise.  Drawn only when needed: a QTable-only run leaves torch's stream alone.
    if rng != "reference":
        seed = random.getrandbits(63)
    elif not all(is_q):
        seed = int(torch.randint(0, 2 ** 62, ()).item())
    else:
        seed = 0
    batch = engine.RunBatch(config, 1, device=device, dtype=torch.float64, seed=seed)
    q0 = abi.pack_tables(batch.game, [a.table if q else None for a, q in zip(agents, is_q)], numpy.float64)
    mlp0 = None
    if batch.mlp is not None:  # MLP agents: the constructor's nn.Linear initialisation (agents.py:137-138)
        mlp0 = numpy.zeros((1, batch.game.mlp_stride), numpy.float32)
        for i, a in enumerate(agents):
            if not is_q[i]:
                flat = numpy.concatenate([v.numpy().reshape(-1) for v in a.state_dict().values()])
                mlp0[0, batch.game.agent[i].mlp_offset:batch.game.agent[i].mlp_offset + flat.size] = flat
    batch.load_state(q0, [[a.epsilon if q else 0.0 for a, q in zip(agents, is_q)]], [float(state[0])], mlp=mlp0)
    chunk = int(chunk_epochs or max(1, min(epochs, print_freq, max(1, 2_000_000 // max(1, max_steps * n)))))
    e0 = 0
    while e0 < epochs:
        E = min(chunk, epochs - e0)
        if rng == "reference":
            for a, x, q in zip(agents, batch.eps[0].tolist(), is_q):
                if q:
                    a.epsilon = x
            u, ra, new_a = reference_streams(agents, environment, E)
            u = numpy.nan_to_num(u, nan=2.0)  # MLP agents: no host draw; ra = -1 leaves their sample to the device
            noisy = environment.noise_prob > 0
            out = batch.scan(E, rng_mode=abi.THRL_RNG_REPLAY_DRAWS, replay_u=u[None], replay_ra=ra[None],
                             replay_new_a=new_a[None] if noisy else None, n_log_runs=1)
        else:
            out = batch.scan(E, n_log_runs=1)
        rewards_log[e0:e0 + E] = out.rewards_log[0].cpu().numpy()
        actions_log[e0:e0 + E] = out.actions_log[0].cpu().numpy()
        eps_now = batch.eps[0].tolist()
        for e in range(e0, e0 + E):
            if not (e + 1) % print_freq:
                _print_progress(config, config["agents"], rewards_log, actions_log, [x for x, q in zip(eps_now, is_q) if q], e, print_freq,
                                time.time() - t, print_eps)
                t = time.time()
        e0 += E

    tabs, cnts, sds = batch.tables(), batch.counters(), batch.mlp_state_dicts(0) if batch.mlp is not None else [None] * n
    for i, a in enumerate(agents):
        if is_q[i]:
            a.table = tabs[i][0].cpu().numpy().astype(numpy.float64)
            a.counter = cnts[i][0].cpu().numpy().view(numpy.uint32).astype(numpy.float64)
            a.epsilon = batch.eps[0, i].item()
        else:
            a.load_state_dict(sds[i])
    environment.state = batch.price[0].item()
    save_run(exp_path, config, agents, rewards_log, actions_log)


class TrainResult:
    """What train_many returns: the device batch (tables, counters, epsilon, price) plus host copies of the logs."""

    def __init__(self, batch, rewards_log, actions_log, stats):
        self.batch, self.rewards_log, self.actions_log, self.stats = batch, rewards_log, actions_log, stats

    curve_hist = None  # [epochs, bins] int64: histogram over all runs of the EWM-smoothed total reward (quantile_bins > 0)

    def quantile_curves(self, qs=(0.5, 0.75, 0.25)):
        """Per-epoch quantiles over the runs of the smoothed total reward: what th_rl/utils.py:141-143 plots as
        median / 75th / 25th, to the histogram's bin width."""
        from . import stats as curve_stats
        if self.curve_hist is None:
            raise ValueError("train_many was called without quantile_bins")
        return curve_stats.quantiles_from_hist(self.curve_hist, qs, *self.curve_range)

    def mean_curves(self):
        """Cross-run mean / std of the per-epoch mean reward and action, from the exact fixed-point sums."""
        R = float(self.n_runs_total)
        s = self.stats.astype(numpy.float64)
        mean_r, mean_x = s[..., 0] / abi.THRL_STATS_SCALE_SUM / R, s[..., 2] / abi.THRL_STATS_SCALE_SUM / R
        var_r = numpy.maximum(s[..., 1] / abi.THRL_STATS_SCALE_SQ / R - mean_r ** 2, 0.0)
        var_x = numpy.maximum(s[..., 3] / abi.THRL_STATS_SCALE_SQ / R - mean_x ** 2, 0.0)
        return mean_r, numpy.sqrt(var_r), mean_x, numpy.sqrt(var_x)


def shard_bounds(total_runs, rank, world):
    """GPU `rank` of `world` owns global runs [lo, hi)."""
    return total_runs * rank // world, total_runs * (rank + 1) // world


def train_many(config, runs, epochs=None, *, seed=0, dtype=None, device=None, log_runs=0, hp=None, chunk_epochs=None,
               export_dir=None, export_runs=0, process_group=None, quantile_bins=0, halflife=1000.0, checkpoint_to=None,
               resume_from=None):
    """`runs` independent runs of one config as a batched device scan (the reference plays them one after another,
    th_rl/main.py:19-21).  Under torch.distributed each rank owns a contiguous shard of the global run ids; results do not
    depend on the sharding because Philox counters use global ids, and the per-epoch statistics are exact integer sums that
    are all-reduced (NCCL on GPUs).  Returns a TrainResult; with export_dir, the first `export_runs` runs are also written
    in the reference's runs/<cfg>/<i>/ layout.
    quantile_bins > 0: also the per-epoch histogram over ALL runs of the EWM-smoothed total reward (th_rl/utils.py:132-145,
    stats.CurveHistogram) -> TrainResult.curve_hist / quantile_curves().
    checkpoint_to: this rank's shard is written as one packed file (checkpoint.py; `{rank}` in the name is substituted) when the
    epochs are done; resume_from: start from such a file instead of a fresh initial state -- `epochs` more epochs are played and
    the result equals an uninterrupted run bit for bit."""
    import torch
    import torch.distributed as dist
    from . import engine

    config = _load_config(config)
    epochs = int(config.get("training", {}).get("epochs", 0) if epochs is None else epochs)
    world = dist.get_world_size(process_group) if dist.is_initialized() else 1
    rank = dist.get_rank(process_group) if dist.is_initialized() else 0
    lo, hi = shard_bounds(int(runs), rank, world)
    if device is None:
        device = "cuda:%d" % (int(os.environ.get("LOCAL_RANK", "0")) if world > 1 else 0)
    dtype = dtype or torch.float32
    hp_local = None if hp is None else numpy.asarray(hp, numpy.float64)[lo:hi]
    from . import stats as curve_stats
    ewm_extra = None
    if resume_from is not None:
        batch, extra, extra_arrays = engine.load_checkpoint(str(resume_from).format(rank=rank), device=device)
        if batch.n_runs != hi - lo or batch.run_id0 != lo:
            raise ValueError("checkpoint holds runs [%d, %d), this rank owns [%d, %d)" % (batch.run_id0, batch.run_id0 + batch.n_runs, lo, hi))
        ewm_extra = dict(extra["curve_hist"], num=extra_arrays["curve_num"]) if "curve_hist" in extra else None
    else:
        batch = engine.RunBatch(config, hi - lo, device=device, dtype=dtype, seed=seed, run_id0=lo, hp=hp_local).init_device()
    n = batch.game.n_agents
    R = hi - lo
    n_log = max(0, min(int(max(log_runs, export_runs)) - lo, R))
    rl = numpy.zeros((n_log, epochs, n))
    al = numpy.zeros((n_log, epochs, n))
    stats = torch.zeros((epochs, n, abi.THRL_STATS_K), dtype=torch.int64, device=batch.device)
    ch = hist = None
    if quantile_bins:
        ch = curve_stats.CurveHistogram(config, R, batch.device, halflife=halflife, bins=quantile_bins)
        hist = torch.zeros((epochs, ch.bins), dtype=torch.int64, device=batch.device)
        if ewm_extra is not None:  # resume the smoothing where the checkpoint left it
            ch.num.copy_(torch.from_numpy(numpy.asarray(ewm_extra["num"], numpy.float64)))
            ch.den_last, ch.epoch = float(ewm_extra["den_last"]), int(ewm_extra["epoch"])
    chunk = int(chunk_epochs or epochs or 1)
    if ch is not None:  # every run's log row lives on the device for one chunk: keep it below ~2 GB
        chunk = max(1, min(chunk, (1 << 31) // max(1, R * n * 16)))
    e0 = 0
    while e0 < epochs:
        E = min(chunk, epochs - e0)
        out = batch.scan(E, n_log_runs=R if ch is not None else n_log, stats=stats[e0:e0 + E])
        if n_log:
            rl[:, e0:e0 + E] = out.rewards_log[:n_log].cpu().numpy()
            al[:, e0:e0 + E] = out.actions_log[:n_log].cpu().numpy()
        if ch is not None:
            ch.update(out.rewards_log, hist[e0:e0 + E])
        e0 += E
    if world > 1:
        dist.all_reduce(stats, group=process_group)
        if hist is not None:
            dist.all_reduce(hist, group=process_group)
    if checkpoint_to is not None:
        extra, extra_arrays = {}, {}
        if ch is not None:
            extra["curve_hist"] = dict(den_last=ch.den_last, epoch=ch.epoch, halflife=ch.halflife)
            extra_arrays["curve_num"] = ch.num.cpu().numpy()
        engine.save_checkpoint(batch, str(checkpoint_to).format(rank=rank), extra=extra, extra_arrays=extra_arrays)
    res = TrainResult(batch, rl, al, stats.cpu().numpy())
    res.n_runs_total = int(runs)
    if ch is not None:
        res.curve_hist, res.curve_range, res.halflife = hist.cpu().numpy(), (ch.lo, ch.hi), ch.halflife
    if export_dir is not None:
        export_runs_to(export_dir, config, res, min(int(export_runs), hi) - lo, first_index=lo)
    return res


def export_runs_to(cfg_dir, config, res, count, first_index=0):
    """Write local runs 0..count-1 as <cfg_dir>/<global index>/{k.npy, k_counter.npy, config.json, log.csv}."""
    b = res.batch
    if count <= 0:
        return
    if not os.path.exists(cfg_dir):
        os.makedirs(cfg_dir)
    import torch
    tabs = [None if t is None else t[:count].cpu().numpy().astype(numpy.float64) for t in b.tables()]
    cnts = [None if c is None else c[:count].cpu().numpy().view(numpy.uint32).astype(numpy.float64) for c in b.counters()]
    state, tstate = numpy.random.get_state(), torch.get_rng_state()
    for r in range(count):
        agents = [AGENTS[a["name"]](**a) for a in config["agents"]]
        sds = b.mlp_state_dicts(r) if b.mlp is not None else [None] * len(agents)
        for i, a in enumerate(agents):
            if tabs[i] is None:
                a.load_state_dict(sds[i])
            else:
                a.table, a.counter = tabs[i][r], cnts[i][r]
        save_run(os.path.join(cfg_dir, str(first_index + r)), config, agents, res.rewards_log[r], res.actions_log[r])
    numpy.random.set_state(state)  # building the carrier objects must not disturb the caller's random streams
    torch.set_rng_state(tstate)
