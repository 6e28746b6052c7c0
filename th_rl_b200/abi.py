"""ctypes mirror of include/thrl.h (the C ABI of the hot path) plus config -> ThrlGame translation.

Nothing here computes anything: it only describes memory layouts.  Field names, order and defaults follow
include/thrl.h, which in turn cites the reference lines each field comes from
(QTable.__init__ th_rl/agents.py:13-27, NoisyPriceState.__init__ th_rl/environments.py:5-13).
"""
import ctypes as C

THRL_ABI_VERSION = 5
THRL_MAX_AGENTS = 16
THRL_MAX_ACTIONS = 255
THRL_STATS_K = 4
THRL_STATS_SCALE_SUM = 4294967296.0
THRL_STATS_SCALE_SQ = 16777216.0

THRL_OK = 0
THRL_ERR_BAD_CONFIG = -1
THRL_ERR_BAD_ARGS = -2
THRL_ERR_UNSUPPORTED = -3
THRL_ERR_CUDA = -4
THRL_ERR_NO_DEVICE = -5

THRL_F32 = 0
THRL_F64 = 1

THRL_AGENT_QTABLE = 0
THRL_AGENT_REINFORCE = 1
THRL_AGENT_ACTORCRITIC = 2
THRL_AGENT_CAC = 3
THRL_MLP_HEADER_WORDS = 4
THRL_PAD_THRESHOLD_BYTES = 58112

THRL_RNG_PHILOX = 0
THRL_RNG_REPLAY_DRAWS = 1
THRL_RNG_REPLAY_ACTIONS = 2


class ThrlAgentSpec(C.Structure):
    _fields_ = [
        ("states", C.c_int32),
        ("actions", C.c_int32),
        ("min_memory", C.c_int32),
        ("capacity", C.c_int32),
        ("action_lo", C.c_double),
        ("action_hi", C.c_double),
        ("max_state", C.c_double),
        ("gamma", C.c_double),
        ("alpha", C.c_double),
        ("eps_end", C.c_double),
        ("eps_step", C.c_double),
        ("table_offset", C.c_int64),
        ("kind", C.c_int32),
        ("hidden", C.c_int32),
        ("lr", C.c_double),
        ("entropy", C.c_double),
        ("mlp_offset", C.c_int64),
        ("row_stride", C.c_int32),
        ("reserved_", C.c_int32),
    ]


class ThrlGame(C.Structure):
    _fields_ = [
        ("n_agents", C.c_int32),
        ("max_steps", C.c_int32),
        ("a", C.c_double),
        ("b", C.c_double),
        ("noise_prob", C.c_double),
        ("agent", ThrlAgentSpec * THRL_MAX_AGENTS),
        ("run_stride", C.c_int64),
        ("ring_len", C.c_int32),
        ("regular", C.c_int32),
        ("mlp_stride", C.c_int64),
        ("mlp_buffer_len", C.c_int32 * THRL_MAX_AGENTS),
    ]


class ThrlScanArgs(C.Structure):
    _fields_ = [
        ("game", C.POINTER(ThrlGame)),
        ("n_runs", C.c_int64),
        ("run_id0", C.c_int64),
        ("epoch_begin", C.c_int32),
        ("epoch_end", C.c_int32),
        ("table_dtype", C.c_int32),
        ("rng_mode", C.c_int32),
        ("seed", C.c_uint64),
        ("q", C.c_void_p),
        ("counter", C.c_void_p),
        ("eps", C.c_void_p),
        ("price", C.c_void_p),
        ("hp", C.c_void_p),
        ("ring", C.c_void_p),
        ("replay_u", C.c_void_p),
        ("replay_ra", C.c_void_p),
        ("replay_new_a", C.c_void_p),
        ("rewards_log", C.c_void_p),
        ("actions_log", C.c_void_p),
        ("n_log_runs", C.c_int64),
        ("stats", C.c_void_p),
        ("trace_actions", C.c_void_p),
        ("trace_rewards", C.c_void_p),
        ("trace_prices", C.c_void_p),
        ("mlp", C.c_void_p),
    ]


# QTable.__init__ defaults (th_rl/agents.py:13-27)
QTABLE_DEFAULTS = dict(states=16, actions=4, action_range=[0, 1], gamma=0.99, buffer="ReplayBuffer", capacity=500,
                       max_state=10, alpha=0.1, eps_end=2e-2, epsilon=0.5, eps_step=5e-4, min_memory=100)
# Reinforce.__init__ defaults (th_rl/agents.py:120-131)
REINFORCE_DEFAULTS = dict(states=4, actions=2, action_range=[0, 1], gamma=0.98, buffer="ReplayBuffer", capacity=50000,
                          min_memory=1000, entropy=0)
# ActorCritic.__init__ defaults (th_rl/agents.py:223-234)
ACTORCRITIC_DEFAULTS = dict(states=4, actions=2, action_range=[0, 1], gamma=0.98, buffer="ReplayBuffer", capacity=50000,
                            min_memory=1000, entropy=0)
# CAC.__init__ defaults (th_rl/agents.py:334-343); it has no `actions`
CAC_DEFAULTS = dict(states=4, action_range=[0, 1], gamma=0.98, buffer="ReplayBuffer", capacity=50000, min_memory=1000, entropy=0)
# NoisyPriceState.__init__ defaults (th_rl/environments.py:5)
ENV_DEFAULTS = dict(action_range=[0, 1], a=10, b=1, max_steps=1, noise_prob=0.05)


def game_from_config(config):
    """Translate the reference's JSON schema (th_rl/some_path/configs/example_config.json) into a ThrlGame.

    Only QTable agents are on this path; unknown keys are ignored exactly as the reference's **kwargs does
    (agents.py:27, environments.py:5).  table_offset / run_stride / ring_len / regular are left for
    thrl_game_layout to fill.
    """
    agents = config["agents"]
    env = dict(ENV_DEFAULTS)
    env.update(config["environment"])
    # trainer.py:21-23
    assert len(agents) == env["nplayers"], "Bad config. Check number of agents."
    if len(agents) > THRL_MAX_AGENTS:
        raise ValueError("at most %d agents per game" % THRL_MAX_AGENTS)
    g = ThrlGame()
    g.n_agents = len(agents)
    g.max_steps = int(env["max_steps"])
    g.a = float(env["a"])
    g.b = float(env["b"])
    g.noise_prob = float(env["noise_prob"])
    for i, ad in enumerate(agents):
        name = ad.get("name", "QTable")
        s = g.agent[i]
        if name in ("Reinforce", "ActorCritic", "CAC"):
            d = dict({"Reinforce": REINFORCE_DEFAULTS, "ActorCritic": ACTORCRITIC_DEFAULTS, "CAC": CAC_DEFAULTS}[name])
            d.update(ad)
            s.kind = {"Reinforce": THRL_AGENT_REINFORCE, "ActorCritic": THRL_AGENT_ACTORCRITIC, "CAC": THRL_AGENT_CAC}[name]
            s.states, s.actions = int(d["states"]), (2 if name == "CAC" else int(d["actions"]))
            s.min_memory, s.capacity = int(d["min_memory"]), int(d["capacity"])
            s.action_lo, s.action_hi = float(d["action_range"][0]), float(d["action_range"][1])
            s.max_state = float("nan")
            s.gamma = float(d["gamma"])
            s.hidden, s.lr, s.entropy = 256, 2e-4, float(d["entropy"])  # agents.py:137-139
            continue
        if name != "QTable":
            raise NotImplementedError(
                "agent %d is %r: the B200 hot path covers the reference's QTable, Reinforce, ActorCritic and CAC agents"
                % (i, name))
        d = dict(QTABLE_DEFAULTS)
        d.update(ad)
        s.states = int(d["states"])
        s.actions = int(d["actions"])
        s.min_memory = int(d["min_memory"])
        s.capacity = int(d["capacity"])
        s.action_lo = float(d["action_range"][0])
        s.action_hi = float(d["action_range"][1])
        s.max_state = float(d["max_state"])
        s.gamma = float(d["gamma"])
        s.alpha = float(d["alpha"])
        s.eps_end = float(d["eps_end"])
        s.eps_step = float(d["eps_step"])
    return g


def eps0_from_config(config):
    return [float(dict(QTABLE_DEFAULTS, **ad)["epsilon"]) if ad.get("name", "QTable") == "QTable" else 0.0
            for ad in config["agents"]]


def mlp_param_count(spec):
    if spec.kind == THRL_AGENT_CAC:
        return 5 * spec.hidden + 3
    p = 2 * spec.hidden + spec.actions * spec.hidden + spec.actions
    return p + spec.hidden + 1 if spec.kind == THRL_AGENT_ACTORCRITIC else p


def mlp_entry_words(spec):
    return 3 if spec.kind == THRL_AGENT_REINFORCE else 4


def mlp_param_shapes(spec):
    """state_dict names -> shapes, in the order the parameters sit in the MLP slab."""
    H, A = spec.hidden, spec.actions
    if spec.kind == THRL_AGENT_CAC:
        return {"fc1.weight": (H, 1), "fc1.bias": (H,), "fc_mu.weight": (1, H), "fc_mu.bias": (1,),
                "fc_std.weight": (1, H), "fc_std.bias": (1,), "fc_v.weight": (1, H), "fc_v.bias": (1,)}
    sh = {"fc1.weight": (H, 1), "fc1.bias": (H,), "fc_pi.weight": (A, H), "fc_pi.bias": (A,)}
    if spec.kind == THRL_AGENT_ACTORCRITIC:
        sh.update({"fc_v.weight": (1, H), "fc_v.bias": (1,)})
    return sh


def mlp_param_names(spec):
    return list(mlp_param_shapes(spec))


# ---- slab layout helpers (include/thrl.h: agent i's table at table_offset_i, rows row_stride_i apart, `actions` columns used)
def table_view(slab, spec):
    """[R, run_stride] slab (numpy array or torch tensor) -> view [R, states+1, actions] of one QTable agent's table."""
    rows = spec.states + 1
    return slab[:, spec.table_offset:spec.table_offset + rows * spec.row_stride].reshape(-1, rows, spec.row_stride)[:, :, :spec.actions]


def pack_tables(game, per_agent, dtype):
    """per_agent: list over agents of arrays [R, states+1, actions] or [states+1, actions] (None for MLP agents) -> numpy slab
    [R, run_stride] in the layout of include/thrl.h (padding cells zero)."""
    import numpy as np
    arrs = [None if a is None else np.asarray(a) for a in per_agent]
    arrs = [a if a is None else (a[None] if a.ndim == 2 else a) for a in arrs]
    R = next((a.shape[0] for a in arrs if a is not None), 1)  # a game may have no Q-tables at all
    out = np.zeros((R, game.run_stride), dtype=dtype)
    for i in range(game.n_agents):
        s = game.agent[i]
        if s.kind == THRL_AGENT_QTABLE:
            table_view(out, s)[...] = arrs[i].astype(dtype)
    return out
