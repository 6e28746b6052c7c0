"""Small invocations of every kernel variant for compute-sanitizer (memcheck / racecheck); exits non-zero on mismatch."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from conftest import load_golden
from test_oracle_golden import golden_inputs
from oracle import oracle
from th_rl_b200 import abi, engine

def replay(name, env, E=3):
    for k in ("THRL_KERNEL", "THRL_LPC_GL"):
        os.environ.pop(k, None)
    os.environ.update(env)
    g = load_golden(name); cfg = g["config"]
    b = engine.RunBatch(cfg, 1, dtype=torch.float64)
    q0, mlp0, u, ra, new_a = golden_inputs(g, b.game, abi.THRL_RNG_REPLAY_DRAWS, np.float64)
    b.load_state(q0, [abi.eps0_from_config(cfg)], [g["p0"]], mlp=mlp0)
    out = b.scan(E, rng_mode=abi.THRL_RNG_REPLAY_DRAWS, replay_u=u[None, :E], replay_ra=ra[None, :E],
                 replay_new_a=None if new_a is None else new_a[None, :E], trace=True, stats=True, n_log_runs=1)
    torch.cuda.synchronize()
    assert np.array_equal(out.trace_actions[0].cpu().numpy(), g["actions"][:E]), name
    print("ok", name, env)

def philox(cfg, R, E, env, dtype=torch.float32):
    for k in ("THRL_KERNEL", "THRL_LPC_GL"):
        os.environ.pop(k, None)
    os.environ.update(env)
    b = engine.RunBatch(cfg, R, dtype=dtype, seed=3).init_device()
    b.scan(E, stats=True)
    a, r = b.greedy_eval(np.full((R, 1), 4.0))
    torch.cuda.synchronize()
    print("ok philox", R, E, env)

replay("c1_example_2q_seed0", {})                       # lut2
replay("c1_example_2q_seed0", {"THRL_KERNEL": "lpc"})   # lane-per-chain
replay("c1_example_2q_seed0", {"THRL_KERNEL": "generic"})
replay("hetero_3q_seed4", {}, E=6)                      # generic, ring across epochs
replay("c4_8q_seed7", {}, E=2)                          # generic, 8 agents, quarter-warp paths
replay("noise_2q_seed3", {}, E=2)
replay("mixed_arq_seed12", {}, E=7)                     # mixed: ActorCritic + Reinforce + QTable
replay("mixed_cc_seed14", {}, E=5)                      # mixed: two CAC agents
cfg = load_golden("c1_example_2q_seed0")["config"]
philox(cfg, 70, 2, {})
philox(cfg, 70, 2, {"THRL_KERNEL": "lpc", "THRL_LPC_GL": "16"})
philox(load_golden("mixed_qa_seed11")["config"], 5, 3, {})
print("all sanitize cases ran")
