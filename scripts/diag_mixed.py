import sys, json
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from conftest import load_golden
from test_oracle_golden import golden_inputs
from oracle import oracle
from th_rl_b200 import abi, engine
name = sys.argv[1] if len(sys.argv) > 1 else "mixed_qr_seed8"
g = load_golden(name); cfg = g["config"]
b = engine.RunBatch(cfg, 1, dtype=torch.float64)
q0, mlp0, u, ra, new_a = golden_inputs(g, b.game, abi.THRL_RNG_REPLAY_DRAWS, np.float64)
E = int(sys.argv[2]) if len(sys.argv) > 2 else g["u"].shape[0]
b.load_state(q0, [abi.eps0_from_config(cfg)], [g["p0"]], mlp=mlp0)
out = b.scan(E, rng_mode=abi.THRL_RNG_REPLAY_DRAWS, replay_u=u[None, :E], replay_ra=ra[None, :E], trace=True)
torch.cuda.synchronize()
ref = oracle.scan(b.game, q0, [abi.eps0_from_config(cfg)], [g["p0"]], E, rng_mode=abi.THRL_RNG_REPLAY_DRAWS, replay_u=u[None, :E], replay_ra=ra[None, :E], mlp=mlp0, trace=True)
got = b.mlp.cpu().numpy()[0]; want = ref.mlp[0]
print("actions equal", np.array_equal(out.trace_actions.cpu().numpy(), ref.trace_actions), "q equal", np.array_equal(b.q.cpu().numpy(), ref.q))
for i in range(b.game.n_agents):
    s = b.game.agent[i]
    if s.kind == 0: continue
    P = abi.mlp_param_count(s); o = s.mlp_offset
    secs = dict(par=(o, o + P), m=(o + P, o + 2 * P), v=(o + 2 * P, o + 3 * P), hdr=(o + 3 * P, o + 3 * P + 4), buf=(o + 3 * P + 4, o + 3 * P + 4 + 3 * b.game.mlp_buffer_len[i]))
    for k, (a, z) in secs.items():
        d = got[a:z].view(np.uint32) != want[a:z].view(np.uint32)
        print(i, k, "mismatching words", int(d.sum()), "of", z - a, "max abs", float(np.nanmax(np.abs(got[a:z] - want[a:z]))) if d.any() else 0.0, "first", (np.flatnonzero(d)[:5]).tolist())
    print("hdr got", got[secs["hdr"][0]:secs["hdr"][1]].view(np.int32), "want", want[secs["hdr"][0]:secs["hdr"][1]].view(np.int32))
