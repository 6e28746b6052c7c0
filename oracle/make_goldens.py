#!/usr/bin/env python
"""Golden-vector generator: runs the UNMODIFIED reference and records its streams.

TEST INFRASTRUCTURE ONLY.  This script needs /root/reference (read-only mount, only
present in the build container); its outputs, tests/golden/*.npz, are committed and are
what travels to the GPU box.  Nothing in the product imports this file.

How it pins the hot path (reference cites, relative to /root/reference):

* th_rl/trainer.py:29-110 `train_one` is called *as is*.  It looks its classes up with
  `eval(name)` in the th_rl.trainer namespace (trainer.py:18,24), so we substitute
  recording subclasses of th_rl.agents.QTable / th_rl.environments.NoisyPriceState under
  the same names; the subclasses only observe, every computation is the parent's.
* The reference never seeds.  We seed python `random`, `numpy.random` and
  torch, and record every draw the path consumes:
    - th_rl/agents.py:81  `random.uniform(0, 1)`   -> u[e,t,i]         (always drawn)
    - th_rl/agents.py:82  `random.choice(space)`   -> ra[e,t,i]        (-1 when not drawn)
    - th_rl/environments.py:28-29 `numpy.random.uniform` -> un[e,t], new_a[e,t]
* Final tables/counters are read back from the files train_one itself wrote
  (agents.py:110-112) and log.csv (trainer.py:107-110), so the artefact layout is pinned too.

Usage:  PYTHONDONTWRITEBYTECODE=1 python oracle/make_goldens.py [--out tests/golden]
"""
import argparse
import contextlib
import io
import json
import os
import random
import sys
import tempfile

import numpy

REFERENCE = os.environ.get("THRL_REFERENCE", "/root/reference")


def _agent(**kw):
    d = dict(name="QTable", gamma=0.95, actions=21, states=100, alpha=0.1, eps_end=0.001,
             epsilon=0.5, eps_step=0.9995, action_range=[0.2, 0.4])
    d.update(kw)
    return d


def _env(**kw):
    d = dict(name="NoisyPriceState", noise_prob=0, a=10, b=1, nplayers=2, max_steps=100)
    d.update(kw)
    return d


# name -> (config, seed).  Sizes are kept small: every case is a few tens of KB compressed.
CASES = {
    # example_config.json hyper-parameters, agent-0 block duplicated (BASELINE C1/C2)
    "c1_example_2q_seed0": (dict(agents=[_agent(), _agent()], environment=_env(),
                                 training=dict(epochs=24, print_freq=1000)), 0),
    "c1_example_2q_seed1": (dict(agents=[_agent(), _agent()], environment=_env(),
                                 training=dict(epochs=12, print_freq=1000)), 1),
    # configs2.json hyper-parameters for the QTable agent (gamma .35, alpha .5, eps .8)
    "configs2_2q_seed2": (dict(agents=[_agent(gamma=0.35, alpha=0.5, epsilon=0.8),
                                       _agent(gamma=0.35, alpha=0.5, epsilon=0.8)],
                               environment=_env(), training=dict(epochs=12, print_freq=1000)), 2),
    # demand-intercept noise on (environments.py:28-31)
    "noise_2q_seed3": (dict(agents=[_agent(), _agent(alpha=0.3)], environment=_env(noise_prob=0.3),
                            training=dict(epochs=10, print_freq=1000)), 3),
    # 3 heterogeneous agents; max_steps < min_memory (batches span episodes), small capacity
    "hetero_3q_seed4": (dict(
        agents=[_agent(states=50, actions=11, action_range=[0.1, 0.3], min_memory=50, capacity=60),
                _agent(states=100, actions=21, action_range=[0.15, 0.3], min_memory=80, capacity=90,
                       gamma=0.8),
                _agent(states=20, actions=5, action_range=[0.05, 0.35], min_memory=37, capacity=500,
                       epsilon=0.9, eps_step=0.9)],
        environment=_env(nplayers=3, max_steps=37), training=dict(epochs=20, print_freq=1000)), 4),
    # max_steps > capacity: only the newest `capacity` transitions are replayed (buffers.py:12)
    "overflow_2q_seed5": (dict(agents=[_agent(capacity=120, min_memory=100),
                                       _agent(capacity=70, min_memory=60)],
                               environment=_env(max_steps=150),
                               training=dict(epochs=8, print_freq=1000)), 5),
    # constructor defaults (agents.py:13-27): states 16, actions 4, range [0,1] -> price hits 0
    "defaults_2q_seed6": (dict(agents=[dict(name="QTable"), dict(name="QTable", actions=6)],
                               environment=_env(max_steps=100, noise_prob=0.05),
                               training=dict(epochs=10, print_freq=1000)), 6),
    # 8 agents, larger action/state space (BASELINE C4 shape, reduced)
    "c4_8q_seed7": (dict(agents=[_agent(states=60, actions=33, action_range=[0.05, 0.15],
                                        alpha=[0.05, 0.1, 0.2, 0.5][i % 4],
                                        gamma=[0.35, 0.8, 0.95, 0.99][i // 2])
                                 for i in range(8)],
                         environment=_env(nplayers=8, max_steps=100),
                         training=dict(epochs=5, print_freq=1000)), 7),
    # the shipped example_config.json pairing: QTable + Reinforce (MLP policy gradient, agents.py:119-219); 20 epochs
    # = two Reinforce updates (min_memory 1000 = every 10 episodes)
    "mixed_qr_seed8": (dict(agents=[_agent(), dict(name="Reinforce", gamma=0.995, actions=21, states=1,
                                                   action_range=[0.2, 0.4])],
                            environment=_env(), training=dict(epochs=20, print_freq=1000)), 8),
    # configs2.json pairing, short episodes, small min_memory: Reinforce batches of 160 transitions spanning 4 episodes
    "mixed_qr_small_seed9": (dict(agents=[_agent(gamma=0.35, alpha=0.5, epsilon=0.8),
                                          dict(name="Reinforce", gamma=0.35, actions=21, states=1,
                                               action_range=[0.2, 0.4], min_memory=150, capacity=400)],
                                  environment=_env(max_steps=40), training=dict(epochs=14, print_freq=1000)), 9),
    # two Reinforce agents and one QTable agent, 7 actions
    "mixed_rqr_seed10": (dict(agents=[dict(name="Reinforce", gamma=0.9, actions=7, states=1, action_range=[0.1, 0.3],
                                           min_memory=100),
                                      _agent(actions=9, states=40, action_range=[0.1, 0.3]),
                                      dict(name="Reinforce", gamma=0.5, actions=5, states=1, action_range=[0.05, 0.2],
                                           min_memory=90, capacity=100)],
                              environment=_env(nplayers=3, max_steps=30), training=dict(epochs=12, print_freq=1000)), 10),
    # ActorCritic (agents.py:222-305): value head with bias 1000, [N,N] advantage broadcast; QTable + ActorCritic
    "mixed_qa_seed11": (dict(agents=[_agent(), dict(name="ActorCritic", gamma=0.98, actions=21, states=1,
                                                    action_range=[0.2, 0.4], min_memory=200)],
                             environment=_env(), training=dict(epochs=8, print_freq=1000)), 11),
    # ActorCritic + Reinforce + QTable, short episodes
    "mixed_arq_seed12": (dict(agents=[dict(name="ActorCritic", gamma=0.9, actions=7, states=1, action_range=[0.1, 0.3],
                                           min_memory=90, capacity=120),
                                      dict(name="Reinforce", gamma=0.5, actions=5, states=1, action_range=[0.05, 0.2],
                                           min_memory=60),
                                      _agent(actions=9, states=40, action_range=[0.1, 0.3])],
                              environment=_env(nplayers=3, max_steps=30), training=dict(epochs=12, print_freq=1000)), 12),
    # CAC (agents.py:333-442): continuous action = sigmoid(Normal(mu, std).sample()), [N,N] log_prob x advantage loss
    "mixed_qc_seed13": (dict(agents=[_agent(), dict(name="CAC", gamma=0.98, states=1, action_range=[0.2, 0.4], min_memory=200)],
                             environment=_env(), training=dict(epochs=8, print_freq=1000)), 13),
    "mixed_cc_seed14": (dict(agents=[dict(name="CAC", gamma=0.9, states=1, action_range=[0.1, 0.3], min_memory=90, capacity=120),
                                     dict(name="CAC", gamma=0.5, states=1, action_range=[0.05, 0.2], min_memory=60)],
                             environment=_env(max_steps=30), training=dict(epochs=12, print_freq=1000)), 14),
    # games made of discrete-action MLP agents only (the BASELINE C5 family): two ActorCritic agents, four updates each
    "mlp_aa_seed15": (dict(agents=[dict(name="ActorCritic", gamma=0.98, actions=21, states=1, action_range=[0.2, 0.4],
                                        min_memory=200),
                                   dict(name="ActorCritic", gamma=0.9, actions=21, states=1, action_range=[0.2, 0.4],
                                        min_memory=150, capacity=300)],
                           environment=_env(), training=dict(epochs=8, print_freq=1000)), 15),
    # two Reinforce agents, short episodes, unequal action grids, ring overflow (capacity 100 < 4 episodes)
    "mlp_rr_seed16": (dict(agents=[dict(name="Reinforce", gamma=0.9, actions=7, states=1, action_range=[0.1, 0.3],
                                        min_memory=100),
                                   dict(name="Reinforce", gamma=0.5, actions=5, states=1, action_range=[0.05, 0.2],
                                        min_memory=90, capacity=100)],
                           environment=_env(max_steps=30), training=dict(epochs=12, print_freq=1000)), 16),
    # Reinforce + two ActorCritic agents
    "mlp_raa_seed17": (dict(agents=[dict(name="Reinforce", gamma=0.8, actions=6, states=1, action_range=[0.1, 0.2],
                                         min_memory=60),
                                    dict(name="ActorCritic", gamma=0.9, actions=7, states=1, action_range=[0.1, 0.3],
                                         min_memory=90, capacity=120),
                                    dict(name="ActorCritic", gamma=0.95, actions=5, states=1, action_range=[0.05, 0.2],
                                         min_memory=45)],
                            environment=_env(nplayers=3, max_steps=30), training=dict(epochs=12, print_freq=1000)), 17),
    # BASELINE C5 at the bench's own shape (bench.py _c5_cfg): two ActorCritic agents with the constructor's min_memory = 1000
    # (agents.py:223-234), i.e. N = 1000-transition batches every 10 episodes; 20 epochs = two updates per agent
    "c5_aa_bench_seed18": (dict(agents=[dict(name="ActorCritic", gamma=0.98, actions=21, states=1, action_range=[0.2, 0.4]),
                                        dict(name="ActorCritic", gamma=0.98, actions=21, states=1, action_range=[0.2, 0.4])],
                                environment=_env(), training=dict(epochs=20, print_freq=1000)), 18),
    # demand noise with MLP agents: the network input is a continuous price (environments.py:28-31 with agents.py:148-163):
    # the shipped pairing QTable + Reinforce, and ActorCritic + CAC
    "noise_qr_seed19": (dict(agents=[_agent(), dict(name="Reinforce", gamma=0.9, actions=21, states=1, action_range=[0.2, 0.4],
                                                    min_memory=200)],
                             environment=_env(noise_prob=0.3), training=dict(epochs=8, print_freq=1000)), 19),
    "noise_ac_seed20": (dict(agents=[dict(name="ActorCritic", gamma=0.9, actions=7, states=1, action_range=[0.1, 0.3],
                                          min_memory=90, capacity=120),
                                     dict(name="CAC", gamma=0.5, states=1, action_range=[0.05, 0.2], min_memory=60)],
                             environment=_env(noise_prob=0.2, max_steps=30), training=dict(epochs=12, print_freq=1000)), 20),
    # entropy regulariser (agents.py:187-189, 298-300, 410-412): Reinforce + ActorCritic, and QTable + CAC with demand noise
    "ent_ra_seed21": (dict(agents=[dict(name="Reinforce", gamma=0.9, actions=7, states=1, action_range=[0.1, 0.3],
                                        min_memory=60, entropy=0.05),
                                   dict(name="ActorCritic", gamma=0.9, actions=5, states=1, action_range=[0.05, 0.2],
                                        min_memory=90, capacity=120, entropy=0.02)],
                           environment=_env(max_steps=30), training=dict(epochs=12, print_freq=1000)), 21),
    "ent_qc_seed22": (dict(agents=[_agent(), dict(name="CAC", gamma=0.98, states=1, action_range=[0.2, 0.4], min_memory=200,
                                                  entropy=0.05)],
                           environment=_env(noise_prob=0.2), training=dict(epochs=8, print_freq=1000)), 22),
}

EVAL_ITERS = 3  # episodes of utils.play_game recorded per case (noise-free cases only)


def _stub_plotly():
    """th_rl/utils.py imports plotly at module level (utils.py:7-9); plotly is not installed here and play_game /
    load_experiment never touch it, so empty stand-in modules are enough to import the file unmodified."""
    import types
    for name in ("plotly", "plotly.graph_objects", "plotly.subplots", "plotly.express"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["plotly.subplots"].make_subplots = None
    sys.modules["plotly"].graph_objects = sys.modules["plotly.graph_objects"]
    sys.modules["plotly"].subplots = sys.modules["plotly.subplots"]
    sys.modules["plotly"].express = sys.modules["plotly.express"]


def record_play_game(out, cfg, seed):
    """th_rl/utils.py:27-47 `play_game` (unmodified) on the agents train_one just saved, loaded back by
    utils.load_experiment (utils.py:12-24; two-agent runs) or create_game + agent.load.  The initial price of every
    episode (environment.reset, environments.py:50-53) is recorded so the rollout can be replayed."""
    import torch
    _stub_plotly()
    import th_rl.utils as rutils
    numpy.random.seed(seed + 7919)
    torch.manual_seed(seed + 7919)
    if len(cfg["agents"]) == 2:
        _, agents, environment, _, _ = rutils.load_experiment(out)
    else:
        _, agents, environment = rutils.create_game(os.path.join(out, "config.json"))
        for i, agent in enumerate(agents):
            agent.load(os.path.join(out, str(i)))
    p0s, orig = [], environment.reset

    def reset():
        s = orig()
        p0s.append(float(s[0]))
        return s

    environment.reset = reset
    new_as, orig_step = [], environment.step

    def step(actions):  # the demand intercept env.step draws (environments.py:28-31), re-derived from the generator state
        st = numpy.random.get_state()
        out_ = orig_step(actions)
        after = numpy.random.get_state()
        numpy.random.set_state(st)
        un = numpy.random.uniform(0, 1)
        new_as.append(numpy.random.uniform(environment.a * 0.7, environment.a) if un < environment.noise_prob else float(environment.a))
        numpy.random.set_state(after)
        return out_

    environment.step = step
    acts, rwds = rutils.play_game(agents, environment, iters=EVAL_ITERS)
    return numpy.array(p0s), numpy.asarray(acts, numpy.float64), numpy.asarray(rwds, numpy.float64), numpy.array(new_as)



class _RandomProxy:
    """Stands in for the `random` module as seen from th_rl.agents (agents.py:3,81-82)."""

    def __init__(self, real, sink):
        self._real, self._sink = real, sink

    def uniform(self, a, b):
        v = self._real.uniform(a, b)
        self._sink.append(("u", v))
        return v

    def choice(self, seq):
        v = self._real.choice(seq)
        self._sink.append(("c", int(v)))
        return v

    def __getattr__(self, k):
        return getattr(self._real, k)


def record_case(cfg, seed):
    sys.path.insert(0, REFERENCE)
    sys.dont_write_bytecode = True
    import torch
    import th_rl.agents as ragents
    import th_rl.environments as renv
    import th_rl.trainer as rtrainer

    draws = []          # QTable exploration draws, in consumption order
    rec = dict(acts=[], rewards=[], prices=[], un=[], new_a=[], q0=[], p0=[], eps_trace=[])

    class QTable(ragents.QTable):
        def __init__(self, **kw):
            super().__init__(**kw)
            rec["q0"].append(self.table.copy())

        def sample_action(self, state):
            n0 = len(draws)
            a = super().sample_action(state)
            new = draws[n0:]
            assert new and new[0][0] == "u" and len(new) <= 2
            rec["acts"].append((new[0][1], new[1][1] if len(new) == 2 else -1, int(a)))
            return a

        def train_net(self):
            super().train_net()
            rec["eps_trace"].append(self.epsilon)

    class Reinforce(ragents.Reinforce):
        def __init__(self, **kw):
            super().__init__(**kw)
            rec.setdefault("mlp0", []).append({k: v.detach().numpy().copy() for k, v in self.state_dict().items()})

        def sample_action(self, state):
            a = super().sample_action(state)
            rec["acts"].append((float("nan"), -1, int(a)))   # no python-random draws: the sample comes from torch's generator
            return a

        def train_net(self):
            super().train_net()
            rec["eps_trace"].append(float("nan"))

    class ActorCritic(ragents.ActorCritic):
        def __init__(self, **kw):
            super().__init__(**kw)
            rec.setdefault("mlp0", []).append({k: v.detach().numpy().copy() for k, v in self.state_dict().items()})

        def sample_action(self, state):
            a = super().sample_action(state)
            rec["acts"].append((float("nan"), -1, int(a)))
            return a

        def train_net(self):
            super().train_net()
            rec["eps_trace"].append(float("nan"))

    class CAC(ragents.CAC):
        def __init__(self, **kw):
            super().__init__(**kw)
            rec.setdefault("mlp0", []).append({k: v.detach().numpy().copy() for k, v in self.state_dict().items()})

        def sample_action(self, state):
            a = super().sample_action(state)   # python float holding a float32 value in (0, 1)
            rec["acts"].append((float("nan"), -1, int(numpy.float32(a).view(numpy.int32))))   # stored as its float32 bit pattern
            return a

        def train_net(self):
            super().train_net()
            rec["eps_trace"].append(float("nan"))

    class NoisyPriceState(renv.NoisyPriceState):
        def reset(self):
            s = super().reset()
            rec["p0"].append(float(s[0]))
            return s

        def step(self, actions):
            st = numpy.random.get_state()
            out = super().step(actions)
            # re-derive the draws step() consumed from the saved generator state
            after = numpy.random.get_state()
            numpy.random.set_state(st)
            un = numpy.random.uniform(0, 1)
            na = numpy.random.uniform(self.a * 0.7, self.a) if un < self.noise_prob else float(self.a)
            numpy.random.set_state(after)
            rec["un"].append(un)
            rec["new_a"].append(na)
            rec["prices"].append(float(out[0][0]))
            rec["rewards"].append(numpy.array(out[1], dtype=numpy.float64))
            return out

    proxy = _RandomProxy(random, draws)
    saved = (ragents.random, rtrainer.QTable, rtrainer.NoisyPriceState, rtrainer.Reinforce, rtrainer.ActorCritic, rtrainer.CAC)
    ragents.random = proxy
    rtrainer.QTable = QTable
    rtrainer.Reinforce = Reinforce
    rtrainer.ActorCritic = ActorCritic
    rtrainer.CAC = CAC
    rtrainer.NoisyPriceState = NoisyPriceState
    torch.set_num_threads(1)
    try:
        random.seed(seed)
        numpy.random.seed(seed)
        torch.manual_seed(seed)
        with tempfile.TemporaryDirectory() as tmp:
            cpath = os.path.join(tmp, "cfg.json")
            with open(cpath, "w") as f:
                json.dump(cfg, f)
            out = os.path.join(tmp, "run0")
            with contextlib.redirect_stdout(io.StringIO()):
                rtrainer.train_one(out, cpath)          # the unmodified reference loop
            n = len(cfg["agents"])
            kinds = [a["name"] for a in cfg["agents"]]
            q_final = [numpy.load(os.path.join(out, "%d.npy" % i)) if kinds[i] == "QTable" else None for i in range(n)]
            c_final = [numpy.load(os.path.join(out, "%d_counter.npy" % i)) if kinds[i] == "QTable" else None for i in range(n)]
            mlp_final = [{k: v.numpy().copy() for k, v in torch.load(os.path.join(out, str(i))).items()}
                         if kinds[i] != "QTable" else None for i in range(n)]
            with open(os.path.join(out, "log.csv")) as f:
                header = [f.readline().strip(), f.readline().strip()]
                log = numpy.loadtxt(f, delimiter=",", ndmin=2)
            ragents.random, rtrainer.QTable, rtrainer.NoisyPriceState, rtrainer.Reinforce, rtrainer.ActorCritic, rtrainer.CAC = saved
            # play_game: noise-free games without CAC agents.  (CAC.get_action builds Normal(mu, 0), agents.py:385-389, which
            # torch's argument validation rejects -- the reference itself raises ValueError there, with the pinned torch 1.10 too.)
            evalrec = None
            if all(a["name"] != "CAC" for a in cfg["agents"]):  # with demand noise the intercepts of the steps are recorded too
                evalrec = record_play_game(out, cfg, seed)
    finally:
        ragents.random, rtrainer.QTable, rtrainer.NoisyPriceState, rtrainer.Reinforce, rtrainer.ActorCritic, rtrainer.CAC = saved

    E = cfg["training"]["epochs"]
    T = cfg["environment"]["max_steps"]
    n = len(cfg["agents"])
    acts = numpy.array(rec["acts"], dtype=object).reshape(E, T, n, 3)
    g = dict(
        config=numpy.array(json.dumps(cfg)),
        seed=numpy.int64(seed),
        p0=numpy.float64(rec["p0"][0]),
        u=acts[..., 0].astype(numpy.float64),
        ra=acts[..., 1].astype(numpy.int32),
        actions=acts[..., 2].astype(numpy.int32),
        rewards=numpy.array(rec["rewards"]).reshape(E, T, n),
        prices=numpy.array(rec["prices"]).reshape(E, T),
        un=numpy.array(rec["un"]).reshape(E, T),
        new_a=numpy.array(rec["new_a"]).reshape(E, T),
        eps_trace=numpy.array(rec["eps_trace"]).reshape(E, n),
        rewards_log=log[:, :n].copy(),
        actions_log=log[:, n:].copy(),
        log_header=numpy.array(header),
    )
    assert len(rec["p0"]) == 1
    if evalrec is not None:  # utils.play_game on the trained agents: [iters] initial prices, [iters*T, n] scaled actions / rewards
        g["eval_p0"], g["eval_actions"], g["eval_rewards"] = evalrec[:3]
        if cfg["environment"].get("noise_prob", 0.05) > 0:
            g["eval_new_a"] = evalrec[3].reshape(EVAL_ITERS, T)  # demand intercept of every step of every episode
    qi = mi = 0
    for i in range(n):
        if kinds[i] == "QTable":
            g["q0_%d" % i] = rec["q0"][qi]; qi += 1
            g["q_final_%d" % i] = q_final[i]
            g["counter_final_%d" % i] = c_final[i]
        else:
            for k, v in rec["mlp0"][mi].items():
                g["mlp0_%d_%s" % (i, k)] = v
            for k, v in mlp_final[i].items():
                g["mlp_final_%d_%s" % (i, k)] = v
            mi += 1
    return g


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(os.path.dirname(__file__), "..", "tests", "golden"))
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    for name, (cfg, seed) in CASES.items():
        if args.only and args.only != name:
            continue
        g = record_case(cfg, seed)
        path = os.path.join(args.out, name + ".npz")
        numpy.savez_compressed(path, **g)
        print("%-24s E=%d T=%d n=%d  %6.1f KB" % (name, g["u"].shape[0], g["u"].shape[1], g["u"].shape[2],
                                                 os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
