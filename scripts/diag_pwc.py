"""Interval-table (pwc) kernel vs the order-exact kernel / the oracle: error magnitudes in replay and free-running mode, and a
quick timing of the noisy C5 shape.  `python scripts/diag_pwc.py [golden ...]`"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from conftest import load_golden
from test_oracle_golden import golden_inputs
from oracle import oracle
from th_rl_b200 import abi, engine, _lib

PWC = ["mixed_arq_seed12", "mixed_cc_seed14", "mixed_qc_seed13", "mixed_qr_small_seed9", "mixed_rqr_seed10", "noise_qr_seed19",
       "noise_ac_seed20", "c5_aa_bench_seed18", "mlp_raa_seed17", "mixed_qr_seed8"]


def slab_err(game, a, m):
    out = []
    for i in range(game.n_agents):
        s = game.agent[i]
        if s.kind == abi.THRL_AGENT_QTABLE:
            continue
        P, o = abi.mlp_param_count(s), s.mlp_offset
        dw = np.abs(a[..., o:o+P].astype(np.float64) - m[..., o:o+P])
        dm = np.abs(a[..., o+P:o+2*P].astype(np.float64) - m[..., o+P:o+2*P]) / (np.abs(m[..., o+P:o+2*P]).max(axis=-1, keepdims=True) + 1e-30)
        dv = np.abs(a[..., o+2*P:o+3*P].astype(np.float64) - m[..., o+2*P:o+3*P]) / (np.abs(m[..., o+2*P:o+3*P]).max(axis=-1, keepdims=True) + 1e-30)
        hdr = np.array_equal(a[..., o+3*P:o+3*P+3].view(np.int32), m[..., o+3*P:o+3*P+3].view(np.int32))
        out.append("agent %d kind %d: |dw| max %.3g (frac>1e-6: %.2g)  m rel %.3g  v rel %.3g  hdr_equal %s nan %d" % (
            i, s.kind, dw.max(), (dw > 1e-6 + 1e-6 * np.abs(m[..., o:o+P])).mean(), dm.max(), dv.max(), hdr, int(np.isnan(a[..., o:o+P]).sum())))
    return out


def replay(name):
    g = load_golden(name); cfg = g["config"]
    res = {}
    for kern in ("pwc", "mixed"):
        os.environ["THRL_KERNEL"] = kern
        b = engine.RunBatch(cfg, 1, dtype=torch.float64)
        q0, mlp0, u, ra, new_a = golden_inputs(g, b.game, abi.THRL_RNG_REPLAY_ACTIONS, np.float64)
        b.load_state(q0, [abi.eps0_from_config(cfg)], [g["p0"]], mlp=mlp0)
        E = g["u"].shape[0]
        out = b.scan(E, rng_mode=abi.THRL_RNG_REPLAY_ACTIONS, replay_u=u[None], replay_ra=ra[None],
                     replay_new_a=None if new_a is None else new_a[None], trace=True, n_log_runs=1)
        torch.cuda.synchronize()
        assert _lib.last_kernel() == kern, _lib.last_kernel()
        res[kern] = (b.mlp.cpu().numpy().copy(), out.trace_rewards.cpu().numpy(), out.trace_prices.cpu().numpy(), b.q.cpu().numpy().copy())
    print(name, "REPLAY rewards equal golden:", np.array_equal(res["pwc"][1][0], g["rewards"]), "prices:", np.array_equal(res["pwc"][2][0], g["prices"]),
          "q equal exact kernel:", np.array_equal(res["pwc"][3], res["mixed"][3]))
    for l in slab_err(b.game, res["pwc"][0], res["mixed"][0]):
        print("   ", l)


def free(name, R=24, E=12, seed=1234):
    g = load_golden(name); cfg = g["config"]
    game = oracle.layout(cfg)
    q0, c0, eps0, p0, mlp0 = oracle.init(game, R, seed=seed, run_id0=7, dtype=np.float32, eps0=abi.eps0_from_config(cfg))
    res = {}
    for kern in ("pwc", "mixed"):
        os.environ["THRL_KERNEL"] = kern
        b = engine.RunBatch(cfg, R, seed=seed, run_id0=7)
        b.load_state(q0, eps0, p0, mlp=mlp0)
        out = b.scan(E, n_log_runs=R, stats=True, trace=True)
        torch.cuda.synchronize()
        res[kern] = dict(mlp=b.mlp.cpu().numpy().copy(), act=out.trace_actions.cpu().numpy(), price=out.trace_prices.cpu().numpy(),
                         rew=out.trace_rewards.cpu().numpy(), q=b.q.cpu().numpy().copy())
    a, m = res["pwc"], res["mixed"]
    cac = [i for i in range(game.n_agents) if game.agent[i].kind == abi.THRL_AGENT_CAC]
    disc = [i for i in range(game.n_agents) if i not in cac]
    same_d = (a["act"][..., disc] == m["act"][..., disc]).reshape(R, -1).all(axis=1) if disc else np.ones(R, bool)
    msg = "%s FREE runs with equal discrete traces: %d/%d" % (name, same_d.sum(), R)
    if cac:
        fa, fm = a["act"][..., cac].view(np.float32), m["act"][..., cac].view(np.float32)
        msg += "  CAC action max |d| %.3g (same-runs %.3g)" % (np.abs(fa - fm).max(), np.abs(fa - fm)[same_d].max() if same_d.any() else -1)
    msg += "  price max |d| on same runs %.3g" % (np.abs(a["price"] - m["price"])[same_d].max() if same_d.any() else -1)
    print(msg)
    if same_d.any():
        for l in slab_err(game, a["mlp"][same_d], m["mlp"][same_d]):
            print("   ", l)
        print("    q equal on same runs:", np.array_equal(a["q"][same_d], m["q"][same_d]))


def timing(case, R, E):
    sys.argv = [sys.argv[0]]
    a = dict(name="ActorCritic", gamma=0.98, actions=21, states=1, action_range=[0.2, 0.4])
    q = dict(name="QTable", gamma=0.95, actions=21, states=100, alpha=0.1, eps_end=0.001, epsilon=0.5, eps_step=0.9995, action_range=[0.2, 0.4])
    r = dict(name="Reinforce", gamma=0.995, actions=21, states=1, action_range=[0.2, 0.4])
    c = dict(name="CAC", gamma=0.98, states=1, action_range=[0.2, 0.4])
    agents, noise = {"noisy_aa": ([a, a], 0.05), "noisy_qr": ([q, r], 0.05), "cac": ([c, c], 0.0)}[case]
    cfg = {"agents": [dict(x) for x in agents], "environment": dict(name="NoisyPriceState", noise_prob=noise, a=10, b=1, nplayers=2, max_steps=100),
           "training": dict(print_freq=500, epochs=20)}
    os.environ.pop("THRL_KERNEL", None)
    b = engine.RunBatch(cfg, R, seed=1).init_device()
    b.scan(10); torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(); o = b.scan(E, stats=True); t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1)
    st = o.stats.cpu().numpy()
    print("%s: kernel %s R=%d E=%d  %.1f ms  %.3g agent-steps/s   mean reward/epoch %.4f -> %.4f  nan params %d" % (
        case, _lib.last_kernel(), R, E, ms, R * 2 * E * 100 / ms * 1e3, st[0, 0, 0] / 2**32 / R, st[-1, 0, 0] / 2**32 / R, int(torch.isnan(b.mlp).sum())))


if __name__ == "__main__":
    names = [x for x in sys.argv[1:] if not x.startswith("-")] or PWC
    for nm in names:
        replay(nm)
    for nm in names:
        free(nm)
    if "--time" in sys.argv:
        for case in ("noisy_aa", "noisy_qr", "cac"):
            timing(case, 16384, 40)
