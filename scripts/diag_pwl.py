"""Lattice (pwl) kernel vs the oracle and the torch goldens: error magnitudes, and a quick C5-shape timing."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from conftest import load_golden
from test_oracle_golden import golden_inputs
from oracle import oracle
from th_rl_b200 import abi, engine

def replay(name):
    g = load_golden(name); cfg = g["config"]
    res = {}
    for kern in ("pwl", "mixed"):
        os.environ.pop("THRL_KERNEL", None)
        if kern == "mixed": os.environ["THRL_KERNEL"] = "mixed"
        b = engine.RunBatch(cfg, 1, dtype=torch.float64)
        q0, mlp0, u, ra, new_a = golden_inputs(g, b.game, abi.THRL_RNG_REPLAY_ACTIONS, np.float64)
        b.load_state(q0, [abi.eps0_from_config(cfg)], [g["p0"]], mlp=mlp0)
        E = g["u"].shape[0]
        out = b.scan(E, rng_mode=abi.THRL_RNG_REPLAY_ACTIONS, replay_u=u[None], replay_ra=ra[None], trace=True, n_log_runs=1)
        torch.cuda.synchronize()
        res[kern] = (b.mlp.cpu().numpy().copy(), out.trace_rewards.cpu().numpy(), out.trace_prices.cpu().numpy(), b.mlp_state_dicts(0))
    game = b.game
    print(name, "rewards equal golden:", np.array_equal(res["pwl"][1][0], g["rewards"]), "prices:", np.array_equal(res["pwl"][2][0], g["prices"]))
    for i in range(game.n_agents):
        s = game.agent[i]
        if s.kind == abi.THRL_AGENT_QTABLE:
            continue
        P, o = abi.mlp_param_count(s), s.mlp_offset
        a, m = res["pwl"][0][0], res["mixed"][0][0]
        dw = np.abs(a[o:o+P].astype(np.float64) - m[o:o+P]); 
        dm = np.abs(a[o+P:o+2*P].astype(np.float64) - m[o+P:o+2*P]) / (np.abs(m[o+P:o+2*P]).max() + 1e-30)
        dv = np.abs(a[o+2*P:o+3*P].astype(np.float64) - m[o+2*P:o+3*P]) / (np.abs(m[o+2*P:o+3*P]).max() + 1e-30)
        hdr_a, hdr_m = a[o+3*P:o+3*P+3].view(np.int32), m[o+3*P:o+3*P+3].view(np.int32)
        et = max(np.abs(res["pwl"][3][i][k].numpy().reshape(g["mlp_final_%d_%s" % (i, k)].shape) - g["mlp_final_%d_%s" % (i, k)]).max() for k in res["pwl"][3][i])
        etm = max(np.abs(res["mixed"][3][i][k].numpy().reshape(g["mlp_final_%d_%s" % (i, k)].shape) - g["mlp_final_%d_%s" % (i, k)]).max() for k in res["mixed"][3][i])
        print("  agent %d kind %d: |w - exact| max %.3g  m rel %.3g  v rel %.3g  hdr %s vs %s   |w - torch| pwl %.3g exact %.3g  nan=%d" % (
            i, s.kind, dw.max(), dm.max(), dv.max(), hdr_a, hdr_m, et, etm, int(np.isnan(a[o:o+P]).sum())))

for nm in (sys.argv[1:] or ["mlp_rr_seed16", "mlp_aa_seed15", "mlp_raa_seed17"]):
    replay(nm)
if len(sys.argv) > 1:
    sys.exit(0)

# C5-shape timing, free running
sys.path.insert(0, ".")
import bench
cfg = bench._c5_cfg(20)
R, E = int(os.environ.get("R", "4096")), 20
for kern in ("pwl",) + (("mixed",) if os.environ.get("WITH_MIXED") else ()):
    os.environ.pop("THRL_KERNEL", None)
    if kern == "mixed": os.environ["THRL_KERNEL"] = "mixed"
    b = engine.RunBatch(cfg, R, seed=1).init_device()
    b.scan(E); torch.cuda.synchronize()
    t0 = time.time(); o = b.scan(E, stats=True); torch.cuda.synchronize(); dt = time.time() - t0
    st = o.stats.cpu().numpy()
    print("%s: R=%d E=%d  %.1f ms  %.3g agent-steps/s   mean reward/epoch %.4f -> %.4f  nan params %d" % (
        kern, R, E, dt * 1e3, R * 2 * E * 100 / dt, st[0, 0, 0] / 2**32 / R, st[-1, 0, 0] / 2**32 / R, int(torch.isnan(b.mlp).sum())))
