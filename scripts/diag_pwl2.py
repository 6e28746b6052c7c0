import os, sys
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from conftest import load_golden
from oracle import oracle
from th_rl_b200 import abi, engine
name = sys.argv[1] if len(sys.argv) > 1 else "mlp_aa_seed15"
cfg = load_golden(name)["config"]
R, E, seed, run_id0 = 24, 12, 1234, 7
game = oracle.layout(cfg)
q0, c0, eps0, p0, mlp0 = oracle.init(game, R, seed=seed, run_id0=run_id0, dtype=np.float32, eps0=abi.eps0_from_config(cfg))
for EE in (2, 4, 6, 8, 12):
    ref = oracle.scan(game, q0, eps0, p0, EE, seed=seed, run_id0=run_id0, stats=True, trace=True, n_threads=0, mlp=mlp0)
    b = engine.RunBatch(cfg, R, seed=seed, run_id0=run_id0); b.load_state(q0, eps0, p0, mlp=mlp0)
    o = b.scan(EE, trace=True); torch.cuda.synchronize()
    acts = o.trace_actions.cpu().numpy()
    same = (acts == ref.trace_actions).reshape(R, -1).all(axis=1)
    got = b.mlp.cpu().numpy()
    line = "E=%d same %d/%d" % (EE, same.sum(), R)
    for i in range(game.n_agents):
        s = game.agent[i]; P, off = abi.mlp_param_count(s), s.mlp_offset
        d = np.abs(got[same, off:off+P].astype(np.float64) - ref.mlp[same, off:off+P])
        rel = d / (1e-6 + 1e-6 * np.abs(ref.mlp[same, off:off+P]))
        idx = np.unravel_index(np.argmax(rel), rel.shape)
        line += " | ag%d max|dw| %.3g worst ratio %.3g at param %d (ref %.6g got %.6g) steps %s" % (i, d.max(), rel.max(), idx[1], ref.mlp[same][idx[0], off+idx[1]], got[same][idx[0], off+idx[1]], got[same][idx[0], off+3*P:off+3*P+1].view(np.int32))
    print(line)
