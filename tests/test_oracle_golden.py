"""The oracle (oracle/thrl_oracle.c) against streams recorded from the unmodified reference.

Every quantity the hot path produces must be BIT-EXACT in the reference's own dtype (f64 tables):
chosen actions, rewards, prices, final tables, visit counters, epsilon and both per-epoch logs.
"""
import numpy as np
import pytest

from oracle import oracle
from th_rl_b200 import abi


def is_mlp(cfg, i):
    return cfg["agents"][i].get("name", "QTable") != "QTable"


def golden_inputs(g, game, rng_mode, dtype):
    """Initial state and replay streams of a golden case, shaped for one run.  MLP agents' recorded samples are forced in
    both replay modes (they come from torch's generator, agents.py:160-163)."""
    cfg = g["config"]
    n = game.n_agents
    q0 = oracle.pack_tables(game, [None if is_mlp(cfg, i) else g["q0_%d" % i] for i in range(n)], dtype)
    mlp0 = None
    if game.mlp_stride:
        mlp0 = oracle.pack_mlp(game, [{k: g["mlp0_%d_%s" % (i, k)] for k in abi.mlp_param_names(game.agent[i])}
                                      if is_mlp(cfg, i) else None for i in range(n)])
    ra = (g["ra"] if rng_mode == abi.THRL_RNG_REPLAY_DRAWS else g["actions"]).copy()
    for i in range(n):
        if is_mlp(cfg, i):
            ra[..., i] = g["actions"][..., i]
    u = np.nan_to_num(g["u"], nan=2.0)
    noisy = cfg["environment"].get("noise_prob", 0.05) > 0
    return q0, mlp0, u, ra, (g["new_a"] if noisy else None)


def run_oracle(g, rng_mode, dtype=np.float64):
    cfg = g["config"]
    game = oracle.layout(cfg)
    q0, mlp0, u, ra, new_a = golden_inputs(g, game, rng_mode, dtype)
    E = g["u"].shape[0]
    return game, oracle.scan(game, q0, [abi.eps0_from_config(cfg)], [g["p0"]], E, rng_mode=rng_mode,
                             replay_u=u[None], replay_ra=ra[None], replay_new_a=None if new_a is None else new_a[None],
                             trace=True, stats=True, mlp=mlp0)


MLP_ATOL = 1e-7  # |weight - torch's weight| after the recorded updates (two to four Adam steps of 2e-4 each); measured <= 1.5e-8
MLP_RTOL = 2e-7  # relative part: ActorCritic's fc_v.bias sits at 1000 (agents.py:244), one float32 ulp there is 6e-5


def check_mlp(g, game, res):
    cfg = g["config"]
    got = oracle.unpack_mlp(game, res.mlp)
    for i in range(game.n_agents):
        if not is_mlp(cfg, i):
            continue
        moved = max(np.abs(g["mlp_final_%d_%s" % (i, k)] - g["mlp0_%d_%s" % (i, k)]).max() for k in got[i])
        assert moved > 1e-4, "the golden run must contain at least one update"
        for k, v in got[i].items():
            ref = g["mlp_final_%d_%s" % (i, k)]
            err = np.abs(v.reshape(ref.shape) - ref)
            assert np.all(err <= MLP_ATOL + MLP_RTOL * np.abs(ref)), (i, k, float(err.max()))


def check_bit_exact(g, game, res):
    n = game.n_agents
    assert np.array_equal(res.trace_actions[0], g["actions"])
    assert np.array_equal(res.trace_rewards[0], g["rewards"])
    assert np.array_equal(res.trace_prices[0], g["prices"])
    tabs = oracle.unpack_tables(game, res.q)
    cnts = oracle.unpack_tables(game, res.counter)
    for i in range(n):
        if is_mlp(g["config"], i):
            continue
        assert np.array_equal(tabs[i][0], g["q_final_%d" % i]), "table of agent %d" % i
        assert np.array_equal(cnts[i][0].astype(np.float64), g["counter_final_%d" % i])
    qt = [i for i in range(n) if not is_mlp(g["config"], i)]
    assert np.array_equal(res.eps[0][qt], g["eps_trace"][-1][qt])
    if game.mlp_stride:
        check_mlp(g, game, res)
    # log.csv round-trips through repr(float) -> exact
    assert np.array_equal(res.rewards_log[0], g["rewards_log"])
    assert np.array_equal(res.actions_log[0], g["actions_log"])
    assert res.price[0] == g["prices"][-1, -1]


def test_replay_draws_f64_bit_exact(golden):
    """Exploration draws replayed, greedy actions chosen by the oracle: everything equals the reference."""
    game, res = run_oracle(golden, abi.THRL_RNG_REPLAY_DRAWS)
    check_bit_exact(golden, game, res)


def test_replay_actions_f64_bit_exact(golden):
    game, res = run_oracle(golden, abi.THRL_RNG_REPLAY_ACTIONS)
    check_bit_exact(golden, game, res)


def test_replay_actions_f32_storage_tolerance(golden):
    """fp32 storage + f64 update arithmetic under teacher forcing: rewards/prices exact, Q within 1e-6 relative."""
    game, res = run_oracle(golden, abi.THRL_RNG_REPLAY_ACTIONS, dtype=np.float32)
    assert np.array_equal(res.trace_rewards[0], golden["rewards"])
    assert np.array_equal(res.trace_prices[0], golden["prices"])
    tabs = oracle.unpack_tables(game, res.q)
    for i in range(game.n_agents):
        if is_mlp(golden["config"], i):
            continue
        ref = golden["q_final_%d" % i]
        rel = np.max(np.abs(tabs[i][0].astype(np.float64) - ref) / np.abs(ref))
        assert rel < 1e-6, rel


def test_stats_are_fixed_point_sums(golden):
    game, res = run_oracle(golden, abi.THRL_RNG_REPLAY_ACTIONS)
    want = np.rint(res.rewards_log[0] * abi.THRL_STATS_SCALE_SUM).astype(np.int64)
    assert np.array_equal(res.stats[:, :, 0], want)
    want = np.rint(res.actions_log[0] ** 2 * abi.THRL_STATS_SCALE_SQ).astype(np.int64)
    assert np.array_equal(res.stats[:, :, 3], want)


def test_quantity_sum_follows_the_interpreters_sum():
    """environments.py:27 uses builtin sum(): compensated while the items are exact floats (MLP agents' actions), plain
    once a numpy.float64 (QTable action) is met.  The oracle's restatement must equal the running interpreter's sum()."""
    import random
    import sys
    if sys.version_info < (3, 12):
        pytest.skip("builtin sum() is only compensated from CPython 3.12 on; the goldens were recorded with 3.12")
    rng = random.Random(7)
    differs = 0
    for _ in range(20000):
        n = rng.randint(1, 8)
        vals = [rng.uniform(0.0, 5.0) for _ in range(n)]
        lead = rng.randint(0, n)
        items = [v if i < lead else np.float64(v) for i, v in enumerate(vals)]
        want = float(sum(items))
        assert oracle.py_sum(vals, lead) == want
        naive = 0.0
        for v in vals:
            naive += v
        differs += naive != want
    assert differs > 0  # the compensation is observable, so the test is not vacuous


def final_state(g, game, dtype=np.float64):
    """The agents train_one saved at the end of the golden run, packed for the oracle / the device."""
    cfg, n = g["config"], game.n_agents
    q = oracle.pack_tables(game, [None if is_mlp(cfg, i) else g["q_final_%d" % i] for i in range(n)], dtype)
    mlp = None
    if game.mlp_stride:
        mlp = oracle.pack_mlp(game, [{k: g["mlp_final_%d_%s" % (i, k)] for k in abi.mlp_param_names(game.agent[i])}
                                     if is_mlp(cfg, i) else None for i in range(n)])
    return q, mlp


def test_greedy_eval_matches_play_game(golden):
    """oracle.greedy_eval against th_rl/utils.py:27-47 `play_game`, recorded from the unmodified reference on the agents the
    golden run saved (oracle/make_goldens.py record_play_game): scaled actions and rewards of every step, bit for bit."""
    if "eval_p0" not in golden:
        pytest.skip("no play_game record: a CAC agent (the reference's CAC.get_action raises ValueError: Normal(mu, 0), "
                    "agents.py:385-389)")
    game = oracle.layout(golden["config"])
    q, mlp = final_state(golden, game)
    new_a = golden["eval_new_a"][None] if "eval_new_a" in golden else None  # demand noise: the intercepts env.step drew
    assert (new_a is not None) == (game.noise_prob > 0)
    acts, rews = oracle.greedy_eval(game, q, golden["eval_p0"][None], mlp=mlp, new_a=new_a)
    assert np.array_equal(acts[0], golden["eval_actions"])
    assert np.array_equal(rews[0], golden["eval_rewards"])
