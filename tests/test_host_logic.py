"""Host-side logic that needs no GPU: C-ABI library loads and exports every declared symbol, config translation,
validation / layout, drop-in class surface, reference random-stream generation, sharding, gloo all-reduce of stats."""
import ctypes
import json
import os
import random
import re
import subprocess
import sys

import numpy as np
import pytest

from th_rl_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from th_rl_b200 import _lib
    L = _lib.lib()
    hdr = open(os.path.join(ROOT, "include", "thrl.h")).read()
    names = set(re.findall(r"\b(thrl_[a-z0-9_]+)\s*\(", hdr))
    assert {"thrl_qtable_scan", "thrl_qtable_scan_host", "thrl_qtable_init", "thrl_greedy_eval", "thrl_game_layout",
            "thrl_ring_bytes", "thrl_last_error", "thrl_abi_version", "thrl_launch_count", "thrl_last_kernel",
            "thrl_last_wave_runs"} <= names
    for nm in names:
        assert hasattr(L, nm), nm
    assert L.thrl_abi_version() == abi.THRL_ABI_VERSION


def test_struct_sizes_match_header():
    """The ctypes mirror and the C structs must agree byte for byte (checked with a tiny C program)."""
    src = r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "thrl.h"
    int main(void){ printf("%zu %zu %zu %zu %zu\n", sizeof(ThrlAgentSpec), sizeof(ThrlGame), sizeof(ThrlScanArgs),
        offsetof(ThrlGame, run_stride), offsetof(ThrlScanArgs, trace_prices)); return 0; }'''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        got = [int(x) for x in subprocess.check_output([exe]).split()]
    want = [ctypes.sizeof(abi.ThrlAgentSpec), ctypes.sizeof(abi.ThrlGame), ctypes.sizeof(abi.ThrlScanArgs),
            abi.ThrlGame.run_stride.offset, abi.ThrlScanArgs.trace_prices.offset]
    assert got == want


def _cfg(**env):
    a = dict(name="QTable", gamma=0.95, actions=21, states=100, alpha=0.1, eps_end=0.001, epsilon=0.5, eps_step=0.9995,
             action_range=[0.2, 0.4])
    e = dict(name="NoisyPriceState", noise_prob=0, a=10, b=1, nplayers=2, max_steps=100)
    e.update(env)
    return {"agents": [dict(a), dict(a)], "environment": e, "training": {"epochs": 3, "print_freq": 500}}


def test_layout_matches_oracle_and_validates():
    from th_rl_b200 import _lib
    from oracle import oracle
    for cfg in (_cfg(), _cfg(max_steps=37), _cfg(max_steps=150)):
        cfg["agents"][1].update(states=50, actions=11, capacity=120, min_memory=60)
        g, o = _lib.game_layout(cfg), oracle.layout(cfg)
        assert (g.run_stride, g.ring_len, g.regular) == (o.run_stride, o.ring_len, o.regular)
        assert [g.agent[i].table_offset for i in range(2)] == [o.agent[i].table_offset for i in range(2)]
    bad = _cfg(a=20)  # price can exceed max_state: the reference dies with IndexError on the first encode
    with pytest.raises(ValueError):
        _lib.game_layout(bad)
    bad = _cfg()
    bad["environment"]["nplayers"] = 3  # trainer.py:21-23
    with pytest.raises(AssertionError):
        _lib.game_layout(bad)
    unknown = _cfg()
    unknown["agents"][1]["name"] = "DQN"  # not an agent of the reference
    with pytest.raises(NotImplementedError):
        _lib.game_layout(unknown)
    cac = _cfg()
    cac["agents"][1] = dict(name="CAC", gamma=0.98, states=1, action_range=[0.2, 0.4], min_memory=200)
    g, o = _lib.game_layout(cac), oracle.layout(cac)
    assert g.mlp_stride == 3 * (5 * 256 + 3) + 4 + 4 * 200 == o.mlp_stride and g.agent[1].kind == abi.THRL_AGENT_CAC
    ac = _cfg()
    ac["agents"][1] = dict(name="ActorCritic", gamma=0.98, actions=21, states=1, action_range=[0.2, 0.4], min_memory=200)
    g, o = _lib.game_layout(ac), oracle.layout(ac)
    Pac = 2 * 256 + 21 * 256 + 21 + 256 + 1
    assert g.mlp_stride == 3 * Pac + 4 + 4 * 200 == o.mlp_stride and g.agent[1].kind == abi.THRL_AGENT_ACTORCRITIC
    # the shipped example_config.json pairing: QTable + Reinforce (MLP 1 -> 256 -> 21)
    mixed = _cfg()
    mixed["agents"][1] = dict(name="Reinforce", gamma=0.995, actions=21, states=1, action_range=[0.2, 0.4])
    g, o = _lib.game_layout(mixed), oracle.layout(mixed)
    P = 2 * 256 + 21 * 256 + 21
    assert g.run_stride == 101 * 21 and g.mlp_stride == 3 * P + 4 + 3 * 1000 == o.mlp_stride
    assert g.mlp_buffer_len[1] == 1000 and g.agent[1].kind == abi.THRL_AGENT_REINFORCE and g.regular == 1
    bad = _cfg()
    bad["agents"][1] = dict(name="Reinforce", actions=21, states=4)  # the environment's state is one number
    with pytest.raises(ValueError):
        _lib.game_layout(bad)


def test_dropin_class_surface():
    from th_rl_b200 import agents, environments, trainer
    np.random.seed(3)
    q = agents.QTable(states=100, actions=21, action_range=[0.2, 0.4], gamma=0.95, unknown_key=1)
    assert q.table.shape == (101, 21) and q.counter.shape == (101, 21) and q.table.dtype == np.float64
    np.random.seed(3)
    assert np.array_equal(q.table, 12.5 / (1 - 0.95) + np.random.randn(101, 21))  # agents.py:29
    assert q.scale(20) == 20 / 20.0 * (0.4 - 0.2) + 0.2
    assert q.encode(np.array([3.3])).tolist() == [33]
    assert q.get_action(np.array([3.3])) == int(np.argmax(q.table[33]))
    with pytest.raises(NotImplementedError):
        q.sample_action(None)
    env = environments.NoisyPriceState(nplayers=2, a=10, b=1, max_steps=100, noise_prob=0)
    nash, cartel = env.get_optimal()
    assert abs(nash - 22.2222222) < 1e-6 and cartel == 25.0  # th_rl/utils.py:91-92
    with pytest.raises(NotImplementedError):
        env.step([0.3, 0.3])
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "c.json")
        json.dump(_cfg(), open(p, "w"))
        config, ags, e = trainer.create_game(p)
        assert len(ags) == 2 and e.nplayers == 2 and config["training"]["epochs"] == 3
        q.save(os.path.join(d, "0"))
        q2 = agents.QTable(states=100, actions=21)
        q2.load(os.path.join(d, "0"))
        assert np.array_equal(q2.table, q.table)


def test_reference_streams_reproduce_recorded_draws(golden):
    """trainer.reference_streams consumes python `random` / numpy.random exactly like the reference's loop: seeded like
    the golden run, it regenerates the recorded u / random-action / demand-intercept streams bit for bit."""
    from th_rl_b200 import trainer
    cfg = golden["config"]
    import torch
    seed = int(golden["seed"])
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "c.json")
        json.dump(cfg, open(p, "w"))
        config, ags, env = trainer.create_game(p)
    n = len(ags)
    for i in range(n):
        if cfg["agents"][i]["name"] == "QTable":
            assert np.array_equal(ags[i].table, golden["q0_%d" % i])
        else:  # same nn.Linear construction order => same initial weights under the same torch seed
            for k, v in ags[i].state_dict().items():
                assert np.array_equal(v.numpy(), golden["mlp0_%d_%s" % (i, k)]), k
    state = env.reset()
    assert state[0] == golden["p0"]
    E = golden["u"].shape[0]
    u, ra, new_a = trainer.reference_streams(ags, env, E)
    assert np.array_equal(u, golden["u"], equal_nan=True) and np.array_equal(ra, golden["ra"]) and np.array_equal(new_a, golden["new_a"])


def test_shard_bounds_cover_everything():
    from th_rl_b200.trainer import shard_bounds
    for total in (1, 7, 65536, 1048576):
        for world in (1, 2, 3, 8):
            b = [shard_bounds(total, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == total and all(b[i][1] == b[i + 1][0] for i in range(world - 1))


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from oracle import oracle
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = _cfg()
    game = oracle.layout(cfg)
    from th_rl_b200.trainer import shard_bounds
    R, E = 24, 3
    lo, hi = shard_bounds(R, rank, world)
    q0, c0, e0, p0 = oracle.init(game, hi - lo, seed=5, run_id0=lo, dtype=np.float32, eps0=abi.eps0_from_config(cfg))
    res = oracle.scan(game, q0, e0, p0, E, seed=5, run_id0=lo, stats=True)  # stands in for the device scan of this shard
    stats = torch.from_numpy(res.stats.copy())
    dist.all_reduce(stats)  # the same call bench.py / train_many make over NCCL
    if rank == 0:
        q.put((stats.numpy(), res.q[:2]))
    dist.destroy_process_group()


def test_sharded_stats_allreduce_gloo_world2():
    """N>1 path on CPU: two ranks own run shards (global ids), all-reduce the fixed-point statistics over gloo; the sum
    equals the unsharded statistics exactly."""
    import torch.multiprocessing as mp
    from oracle import oracle
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    stats, qhead = q.get(timeout=120)
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    cfg = _cfg()
    game = oracle.layout(cfg)
    q0, c0, e0, p0 = oracle.init(game, 24, seed=5, dtype=np.float32, eps0=abi.eps0_from_config(cfg))
    whole = oracle.scan(game, q0, e0, p0, 3, seed=5, stats=True)
    assert np.array_equal(stats, whole.stats)
    assert np.array_equal(qhead, whole.q[:2])


def test_c_abi_error_paths_need_no_device():
    """Argument and configuration errors are reported (negative ThrlStatus + message) before any device access, so they can
    be checked on a CPU-only box; a valid call on a box without a GPU fails loudly with THRL_ERR_NO_DEVICE (no fallback)."""
    import ctypes as C
    from th_rl_b200 import _lib
    L = _lib.lib()
    cfg = _cfg()
    g = _lib.game_layout(cfg)
    R, n = 2, 2
    q = np.zeros((R, g.run_stride), np.float32)
    eps = np.full((R, n), 0.5)
    price = np.full((R,), 3.0)

    def args(**kw):
        a = abi.ThrlScanArgs()
        a.game = C.pointer(g)
        a.n_runs, a.epoch_begin, a.epoch_end = R, 0, 1
        a.table_dtype, a.rng_mode = abi.THRL_F32, abi.THRL_RNG_PHILOX
        a.q, a.eps, a.price = q.ctypes.data, eps.ctypes.data, price.ctypes.data
        for k, v in kw.items():
            setattr(a, k, v)
        return a

    def err():
        return L.thrl_last_error().decode()

    assert L.thrl_qtable_scan(None, None) == abi.THRL_ERR_BAD_ARGS
    assert L.thrl_qtable_scan(C.byref(args(table_dtype=7)), None) == abi.THRL_ERR_BAD_ARGS and "table_dtype" in err()
    assert L.thrl_qtable_scan(C.byref(args(rng_mode=9)), None) == abi.THRL_ERR_BAD_ARGS
    assert L.thrl_qtable_scan(C.byref(args(eps=None)), None) == abi.THRL_ERR_BAD_ARGS
    assert L.thrl_qtable_scan(C.byref(args(q=None)), None) == abi.THRL_ERR_BAD_ARGS and "q must not be NULL" in err()
    assert L.thrl_qtable_scan(C.byref(args(rng_mode=abi.THRL_RNG_REPLAY_DRAWS)), None) == abi.THRL_ERR_BAD_ARGS  # no streams
    assert L.thrl_qtable_scan(C.byref(args(n_log_runs=5)), None) == abi.THRL_ERR_BAD_ARGS
    assert L.thrl_qtable_scan(C.byref(args(epoch_end=-1)), None) == abi.THRL_ERR_BAD_ARGS
    bad = abi.game_from_config(_cfg(a=20))  # reference: IndexError on the first encode
    a = args()
    a.game = C.pointer(bad)
    assert L.thrl_qtable_scan(C.byref(a), None) == abi.THRL_ERR_BAD_CONFIG and "IndexError" in err()
    mixed = _cfg()
    mixed["agents"][1] = dict(name="Reinforce", gamma=0.9, actions=5, states=1, action_range=[0.1, 0.2])
    gm = _lib.game_layout(mixed)
    a = args()
    a.game = C.pointer(gm)
    assert L.thrl_qtable_scan(C.byref(a), None) == abi.THRL_ERR_BAD_ARGS and "mlp" in err()
    ent = abi.game_from_config(mixed)
    ent.agent[1].entropy = 0.01  # the entropy regulariser is part of the path (agents.py:187-189): accepted
    assert L.thrl_game_layout(C.byref(ent)) == abi.THRL_OK
    ent.agent[1].entropy = float("nan")
    assert L.thrl_game_layout(C.byref(ent)) == abi.THRL_ERR_BAD_CONFIG and "entropy" in err()
    assert L.thrl_qtable_scan(C.byref(args(n_runs=0)), None) == abi.THRL_OK  # nothing to do
    import torch
    if not torch.cuda.is_available():
        assert L.thrl_qtable_scan(C.byref(args()), None) == abi.THRL_ERR_NO_DEVICE and "no CPU fallback" in err()
        from th_rl_b200 import engine
        with pytest.raises(RuntimeError):
            engine.RunBatch(cfg, 4)


def test_host_pipeline_chunk_bounds():
    """engine.chunk_bounds (the launches scan_from_host cuts a batch into): every run exactly once, half-weight first and last
    chunk, and whole multiples of the kernel's resident-run count once the chunks span several rounds of the persistent grid."""
    from th_rl_b200.engine import chunk_bounds
    for R, n, wave in ((131072, 12, 148 * 23), (16384, 8, 147 * 16), (4096, 2, 147 * 14), (1000, 7, 0), (5, 12, 3404), (1, 1, 0),
                       (30000, 3, 3404), (100000, 12, 1)):
        b = chunk_bounds(R, n, wave)
        assert b[0] == 0 and b[-1] == R and all(x <= y for x, y in zip(b, b[1:])), (R, n, wave, b)
        assert len(b) == min(n, R) + 1
    b = chunk_bounds(131072, 12, 3404)  # 38.5 rounds: every inner boundary on a round boundary
    assert all(x % 3404 == 0 for x in b[1:-1])
    sizes = [y - x for x, y in zip(b, b[1:])]
    assert sizes[0] <= max(sizes[1:-1]) // 2 + 3404 and max(sizes) <= 4 * 3404
    b = chunk_bounds(16384, 8, 2352)    # 7 rounds for 8 chunks: plain weighted split, no alignment
    assert [y - x for x, y in zip(b, b[1:])][1:-1] == [2340, 2341, 2340, 2341, 2340, 2341] or any(x % 2352 for x in b[1:-1])
    assert chunk_bounds(4096, 2, 2058) == [0, 2048, 4096]


def test_greedy_cache_carry_rules_are_exact():
    """The rule the general kernel uses to keep its greedy-action cache through the update pass (thrl_scan_generic.cuh,
    DESIGN.md 4.2), restated on one row: from (max, first argmax) over all columns but k -- taken before the write, and
    corrected by `patch` for an intervening write to another column -- the first argmax after `row[k] = x` is k when
    x > max or (x == max and k < argmax), else the carried argmax.  Checked against numpy.argmax on random rows with ties."""
    rng = np.random.default_rng(3)

    def patch(rm, ra, ok, kk, x):  # (max, argmax, valid) of the tracked columns after column kk := x
        if kk != ra:
            if x > rm or (x == rm and kk < ra):
                rm, ra = x, kk
        elif x >= rm:
            rm = x
        else:
            ok = False
        return rm, ra, ok

    unknown = 0
    for _ in range(4000):
        A = int(rng.integers(2, 12))
        row = rng.integers(0, 4, A).astype(np.float32)  # few distinct values: many ties
        k = int(rng.integers(A))                         # the column the next transition writes
        others = [c for c in range(A) if c != k]
        rm, ra, ok = float(max(row[others])), int(others[int(np.argmax(row[others]))]), True
        if rng.random() < 0.5:                           # the transition in between writes the same row (state repeats)
            k2, x2 = int(rng.integers(A)), np.float32(rng.integers(0, 4))
            row[k2] = x2
            if k2 != k:
                rm, ra, ok = patch(rm, ra, ok, k2, float(x2))
        x = np.float32(rng.integers(0, 4))
        row[k] = x
        if not ok:
            unknown += 1
            continue
        got = k if (x > rm or (x == rm and k < ra)) else ra
        assert got == int(np.argmax(row)), (row, k, x, rm, ra)
    assert unknown < 1200  # only a lowered maximum under a repeated state loses the entry


def test_hbm_update_model_equals_sequential_train_net():
    """scripts/model_hbm_update.py: the HBM kernel's update (row groups, first-writer tags on the gathered copy, untagged
    max + rewritten-cell lists, exact greedy refresh) == the sequential form of agents.py:59-78 on random batches."""
    import runpy
    runpy.run_path(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts", "model_hbm_update.py"),
                   run_name="__main__")


def test_bench_workloads_are_valid_games():
    """Every workload bench.py measures is a game the library lays out (no GPU needed): catches a typo in a bench config before
    the GPU box does."""
    import bench
    from th_rl_b200 import _lib
    assert set(bench.EXTRAS) <= set(bench.WORKLOADS) and "c2" in bench.WORKLOADS
    for name, wl in bench.WORKLOADS.items():
        g = _lib.game_layout(wl["config"])
        assert g.n_agents == wl["agents"] and g.max_steps == bench.MAX_STEPS, name
        assert wl["runs_per_gpu"] > 0 and wl["epochs"] > 0 and wl["algo_bytes"] > 0, name
        if wl["hp"] is not None:
            assert wl["hp"](7, wl["agents"]).shape == (7, wl["agents"], 4), name
