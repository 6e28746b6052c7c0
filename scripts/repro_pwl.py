import os, sys
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from conftest import load_golden
from oracle import oracle
from th_rl_b200 import abi, engine, _lib
g = load_golden(sys.argv[1]); cfg = g["config"]
game = oracle.layout(cfg)
R, E, stats, trace, nlog, seed = [int(v) for v in sys.argv[2:8]]
q0, c0, eps0, p0, mlp0 = oracle.init(game, R, seed=seed, dtype=np.float32, eps0=abi.eps0_from_config(cfg))
b = engine.RunBatch(cfg, R, seed=seed); b.load_state(q0, eps0, p0, mlp=mlp0)
try:
    b.scan(E, n_log_runs=nlog, trace=bool(trace), stats=bool(stats)); torch.cuda.synchronize()
    print(sys.argv[2:], "ok", _lib.last_kernel(), flush=True)
except Exception as ex:
    print(sys.argv[2:], "FAILED", str(ex)[:80], flush=True)
