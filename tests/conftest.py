import glob
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "lattice_game: the test's game runs on the lattice kernel by default (kernel_choice fixture)")


def golden_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    g["config"] = json.loads(str(g["config"]))
    return g


@pytest.fixture(params=golden_names())
def golden(request):
    g = load_golden(request.param)
    g["name"] = request.param
    return g


def uses_lattice_kernel(cfg):
    """Games the library plays with the lattice kernel (th_rl_b200/csrc/thrl_scan_pwl.cuh; mirrors plan_pwl in thrl.cu):
    Reinforce / ActorCritic agents, optionally next to QTable agents of a regular game, noise-free demand, at most 1024
    joint actions.  Its MLP results match the order-exact kernel / the oracle to float32 rounding instead of bit for bit;
    THRL_KERNEL=mixed selects the order-exact kernel."""
    from oracle import oracle
    from th_rl_b200 import abi
    if cfg["environment"].get("noise_prob", 0.05) > 0:
        return False
    g = oracle.layout(cfg)
    kinds = [g.agent[i].kind for i in range(g.n_agents)]
    mlp = [k in (abi.THRL_AGENT_REINFORCE, abi.THRL_AGENT_ACTORCRITIC) for k in kinds]
    if not any(mlp) or any(k == abi.THRL_AGENT_CAC for k in kinds):
        return False
    if any(g.agent[i].entropy != 0 for i in range(g.n_agents) if mlp[i]):
        return False  # the entropy regulariser runs on the interval-table kernel
    if not all(mlp) and not g.regular:
        return False
    joint = 1
    for i in range(g.n_agents):
        if mlp[i] and g.agent[i].actions > 31:
            return False
        joint *= g.agent[i].actions
    return joint <= 1024


def uses_pwc_kernel(cfg):
    """Games the library plays with the interval-table kernel (th_rl_b200/csrc/thrl_scan_pwc.cuh; mirrors plan_pwc in thrl.cu):
    games with MLP agents that the lattice kernel does not take -- demand noise, CAC agents, QTable agents whose batches span
    episodes -- with at most 256 hidden units and 32 head columns per MLP agent.  Same tolerance statement as the lattice
    kernel; THRL_KERNEL=mixed selects the order-exact kernel, THRL_KERNEL=pwc forces this one on lattice games too."""
    from oracle import oracle
    from th_rl_b200 import abi
    if uses_lattice_kernel(cfg):
        return False
    return pwc_can_play(cfg)


def pwc_can_play(cfg):
    from oracle import oracle
    from th_rl_b200 import abi
    g = oracle.layout(cfg)
    any_mlp = False
    for i in range(g.n_agents):
        s = g.agent[i]
        if s.kind == abi.THRL_AGENT_QTABLE:
            continue
        any_mlp = True
        nc = 3 if s.kind == abi.THRL_AGENT_CAC else s.actions + (1 if s.kind == abi.THRL_AGENT_ACTORCRITIC else 0)
        if s.hidden < 1 or s.hidden > 256 or nc > 32:
            return False
    return any_mlp
