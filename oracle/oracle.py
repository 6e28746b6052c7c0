"""Python handle on oracle/thrl_oracle.c (the CPU restatement of the reference hot path).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package never imports this module.

Parity status: pinned (tests/test_oracle_golden.py vs tests/golden/*.npz recorded from the live reference).
"""
import ctypes as C
import os
import subprocess

import numpy as np

from th_rl_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libthrl_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "thrl_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "thrl.h")
    stale = (not os.path.exists(_SO)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(_SO) for p in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.thrl_oracle_qtable_scan.argtypes = [C.POINTER(abi.ThrlScanArgs), C.c_int]
        _lib.thrl_oracle_qtable_scan.restype = C.c_int
        _lib.thrl_oracle_game_layout.argtypes = [C.POINTER(abi.ThrlGame)]
        _lib.thrl_oracle_game_layout.restype = C.c_int
        _lib.thrl_oracle_qtable_init.argtypes = [C.POINTER(abi.ThrlGame), C.c_int64, C.c_int64, C.c_uint64, C.c_int32,
                                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.thrl_oracle_qtable_init.restype = C.c_int
        _lib.thrl_oracle_game_init.argtypes = _lib.thrl_oracle_qtable_init.argtypes + [C.c_void_p]
        _lib.thrl_oracle_game_init.restype = C.c_int
        _lib.thrl_oracle_greedy_eval.argtypes = [C.POINTER(abi.ThrlGame), C.c_int64, C.c_int32, C.c_void_p, C.c_int32,
                                                 C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.thrl_oracle_greedy_eval.restype = C.c_int
        _lib.thrl_oracle_greedy_eval_mlp.argtypes = [C.POINTER(abi.ThrlGame), C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                                     C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.thrl_oracle_greedy_eval_mlp.restype = C.c_int
        _lib.thrl_oracle_greedy_eval_noise.argtypes = [C.POINTER(abi.ThrlGame), C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.thrl_oracle_greedy_eval_noise.restype = C.c_int
        _lib.thrl_oracle_online_cores.restype = C.c_int
        _lib.thrl_oracle_py_sum.restype = C.c_double
        _lib.thrl_oracle_py_sum.argtypes = [C.POINTER(C.c_double), C.c_int, C.c_int]
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def layout(config):
    g = abi.game_from_config(config)
    rc = lib().thrl_oracle_game_layout(C.byref(g))
    if rc != 0:
        raise ValueError("bad config (oracle layout rc=%d)" % rc)
    return g


def pack_tables(game, per_agent, dtype):
    """per_agent: list over agents of arrays [R, states+1, actions] (or [states+1, actions]; None for MLP agents)
    -> slab [R, run_stride] (include/thrl.h layout; padding cells zero)."""
    return abi.pack_tables(game, [None if a is None else np.asarray(a, dtype=np.float64) for a in per_agent], dtype)


def pack_mlp(game, per_agent, R=1):
    """per_agent: list over agents of state_dict-like mappings (fc1.weight, fc1.bias, fc_pi.weight, fc_pi.bias; None for
    QTable agents), shared by all R runs -> MLP slab [R, mlp_stride] float32 (Adam state, header and buffer zeroed)."""
    slab = np.zeros((R, max(1, game.mlp_stride)), np.float32)
    for i in range(game.n_agents):
        s = game.agent[i]
        if s.kind == abi.THRL_AGENT_QTABLE:
            continue
        d = per_agent[i]
        flat = np.concatenate([np.asarray(d[k], np.float32).reshape(-1) for k in abi.mlp_param_names(s)])
        assert flat.size == abi.mlp_param_count(s)
        slab[:, s.mlp_offset:s.mlp_offset + flat.size] = flat
    return slab


def unpack_mlp(game, slab, run=0):
    """-> list over agents of {name: array} in torch state_dict shapes (None for QTable agents)."""
    out = []
    for i in range(game.n_agents):
        s = game.agent[i]
        if s.kind == abi.THRL_AGENT_QTABLE:
            out.append(None)
            continue
        p = np.asarray(slab)[run, s.mlp_offset:s.mlp_offset + abi.mlp_param_count(s)]
        d, o = {}, 0
        for k, shp in abi.mlp_param_shapes(s).items():
            cnt = int(np.prod(shp))
            d[k] = p[o:o + cnt].reshape(shp).copy()
            o += cnt
        out.append(d)
    return out


def unpack_tables(game, slab):
    return [abi.table_view(slab, game.agent[i]) if game.agent[i].kind == abi.THRL_AGENT_QTABLE else None
            for i in range(game.n_agents)]


class ScanResult:
    pass


def scan(game, q, eps, price, epochs, *, epoch_begin=0, counter=None, hp=None, rng_mode=abi.THRL_RNG_PHILOX, seed=0,
         run_id0=0, replay_u=None, replay_ra=None, replay_new_a=None, n_log_runs=None, stats=False, trace=False,
         n_threads=1, mlp=None):
    """Plays epochs [epoch_begin, epoch_begin+epochs) for every run.  q [R, run_stride] f32/f64, eps [R, n], price [R]
    are copied, the copies are advanced and returned in a ScanResult."""
    n, T = game.n_agents, game.max_steps
    q = np.ascontiguousarray(q).copy()
    R = q.shape[0]
    dtype = abi.THRL_F64 if q.dtype == np.float64 else abi.THRL_F32
    assert q.dtype in (np.float32, np.float64) and q.shape == (R, max(1, game.run_stride)) or q.shape == (R, game.run_stride)
    eps = np.ascontiguousarray(eps, dtype=np.float64).reshape(R, n).copy()
    price = np.ascontiguousarray(price, dtype=np.float64).reshape(R).copy()
    counter = np.zeros((R, game.run_stride), np.uint32) if counter is None else np.ascontiguousarray(counter, np.uint32).copy()
    E = int(epochs)
    n_log = R if n_log_runs is None else int(n_log_runs)
    res = ScanResult()
    res.q, res.eps, res.price, res.counter = q, eps, price, counter
    res.mlp = None if mlp is None else np.ascontiguousarray(mlp, np.float32).reshape(R, -1).copy()
    assert (game.mlp_stride == 0) == (mlp is None), "games with MLP agents need the mlp slab"
    res.rewards_log = np.zeros((n_log, E, n), np.float64)
    res.actions_log = np.zeros((n_log, E, n), np.float64)
    res.stats = np.zeros((E, n, abi.THRL_STATS_K), np.int64) if stats else None
    res.trace_actions = np.zeros((R, E, T, n), np.int32) if trace else None
    res.trace_rewards = np.zeros((R, E, T, n), np.float64) if trace else None
    res.trace_prices = np.zeros((R, E, T), np.float64) if trace else None
    keep = []

    def inp(a, dt, shape):
        if a is None:
            return None
        a = np.ascontiguousarray(a, dtype=dt).reshape(shape)
        keep.append(a)
        return a

    a = abi.ThrlScanArgs()
    a.game = C.pointer(game)
    a.n_runs, a.run_id0 = R, run_id0
    a.epoch_begin, a.epoch_end = epoch_begin, epoch_begin + E
    a.table_dtype, a.rng_mode, a.seed = dtype, rng_mode, seed
    a.q, a.counter, a.eps, a.price = _ptr(q), _ptr(counter), _ptr(eps), _ptr(price)
    a.hp = _ptr(inp(hp, np.float64, (R, n, 4)))
    a.replay_u = _ptr(inp(replay_u, np.float64, (R, E, T, n)))
    a.replay_ra = _ptr(inp(replay_ra, np.int32, (R, E, T, n)))
    a.replay_new_a = _ptr(inp(replay_new_a, np.float64, (R, E, T)))
    a.rewards_log, a.actions_log, a.n_log_runs = _ptr(res.rewards_log), _ptr(res.actions_log), n_log
    a.stats = _ptr(res.stats)
    a.trace_actions, a.trace_rewards, a.trace_prices = _ptr(res.trace_actions), _ptr(res.trace_rewards), _ptr(res.trace_prices)
    a.mlp = _ptr(res.mlp)
    rc = lib().thrl_oracle_qtable_scan(C.byref(a), n_threads)
    if rc != 0:
        raise IndexError("oracle: table row/action out of range (rc=%d)" % rc)
    return res


def init(game, n_runs, *, seed=0, run_id0=0, dtype=np.float32, hp=None, eps0=None):
    n = game.n_agents
    q = np.zeros((n_runs, game.run_stride), dtype)
    counter = np.zeros((n_runs, game.run_stride), np.uint32)
    eps = np.zeros((n_runs, n), np.float64)
    price = np.zeros((n_runs,), np.float64)
    eps0 = np.ascontiguousarray(eps0 if eps0 is not None else [0.5] * n, np.float64)
    hp_a = None if hp is None else np.ascontiguousarray(hp, np.float64)
    mlp = np.zeros((n_runs, game.mlp_stride), np.float32) if game.mlp_stride else None
    rc = lib().thrl_oracle_game_init(C.byref(game), n_runs, run_id0, seed,
                                     abi.THRL_F64 if dtype == np.float64 else abi.THRL_F32,
                                     _ptr(hp_a), _ptr(eps0), _ptr(q), _ptr(counter), _ptr(eps), _ptr(price), _ptr(mlp))
    assert rc == 0
    if game.mlp_stride:
        return q, counter, eps, price, mlp
    return q, counter, eps, price


def greedy_eval(game, q, price0, mlp=None, new_a=None):
    q = np.ascontiguousarray(q)
    R = q.shape[0]
    mlp = None if mlp is None else np.ascontiguousarray(mlp, np.float32)
    price0 = np.ascontiguousarray(price0, np.float64).reshape(R, -1)
    iters = price0.shape[1]
    n, T = game.n_agents, game.max_steps
    new_a = None if new_a is None else np.ascontiguousarray(new_a, np.float64).reshape(R, iters, T)
    rewards = np.zeros((R, iters * T, n))
    actions = np.zeros((R, iters * T, n))
    rc = lib().thrl_oracle_greedy_eval_noise(C.byref(game), R, abi.THRL_F64 if q.dtype == np.float64 else abi.THRL_F32,
                                             _ptr(q), _ptr(mlp), iters, _ptr(price0), _ptr(new_a), _ptr(rewards), _ptr(actions))
    if rc != 0:
        raise IndexError("oracle greedy_eval rc=%d" % rc)
    return actions, rewards


def py_sum(items, lead):
    """The restated CPython >= 3.12 builtin sum() over scaled quantities whose first `lead` items are exact floats."""
    arr = (C.c_double * len(items))(*[float(v) for v in items])
    return lib().thrl_oracle_py_sum(arr, len(items), int(lead))


def online_cores():
    return lib().thrl_oracle_online_cores()


def curve_hist(rewards_log, decay, den, num, lo, hi, n_bins):
    """CPU restatement of thrl_curve_hist (include/thrl.h; th_rl/utils.py:136-145): rewards_log [R, E, n] -> (hist [E, n_bins]
    int64, num [R]) with exactly the device's operations: x = ((0 + r_0) + r_1) + ..., num = num * decay + x, v = num / den_t,
    bin = floor((v - lo) * (n_bins / (hi - lo))) clamped."""
    rewards_log = np.asarray(rewards_log, np.float64)
    R, E, n = rewards_log.shape
    num = np.array(num, np.float64).copy()
    hist = np.zeros((E, n_bins), np.int64)
    inv_width = float(n_bins) / (hi - lo)
    for e in range(E):
        x = np.zeros(R)
        for i in range(n):
            x = x + rewards_log[:, e, i]
        num = num * decay + x
        v = num / den[e]
        b = np.floor((v - lo) * inv_width)
        b = np.where(np.isnan(b), n_bins - 1, np.clip(b, 0, n_bins - 1)).astype(np.int64)
        np.add.at(hist[e], b, 1)
    return hist, num
