"""Free-running (Philox) mode against the reference, distributionally.

The device's counter-based streams cannot be bit-compared with the reference's MT19937 / torch generators, so
replay parity (tests/test_gpu_parity.py) is complemented here: tests/golden/ensemble/*.npz holds, for 32 seeded runs of the
UNMODIFIED reference (oracle/make_ensemble.py: th_rl/trainer.py:29-110 train_one, 2,000 epochs), the 50-epoch window means of
log.csv's rewards and actions columns.  The same config is trained on a few hundred free-running runs here and the two
ensembles are compared window by window (difference of means against its standard error) and at the end of training
(two-sample Kolmogorov-Smirnov).  Cases: 2 x QTable (BASELINE C1/C2 hyper-parameters) and the shipped QTable + Reinforce pairing.
"""
import json
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
Z_MAX = 4.5        # |difference of window means| / standard error; 160 comparisons per case
KS_P_MIN = 1e-3


def _fixture(name):
    g = np.load(os.path.join(ROOT, "tests", "golden", "ensemble", name + ".npz"))
    return json.loads(str(g["config"])), int(g["window"]), g["rewards"], g["actions"]


def _windows(log, window):
    R, E, n = log.shape
    return log.reshape(R, E // window, window, n).mean(axis=2)


def _compare(ref, dev, what):
    """ref [Rr, W, n], dev [Rd, W, n]: window means of both ensembles."""
    from scipy import stats as sst
    mr, md = ref.mean(0), dev.mean(0)
    se = np.sqrt(ref.var(0, ddof=1) / ref.shape[0] + dev.var(0, ddof=1) / dev.shape[0])
    z = np.abs(mr - md) / np.maximum(se, 1e-12)
    assert z.max() < Z_MAX, "%s: window %s differs by %.1f standard errors (ref %.4f, device %.4f)" % (
        what, np.unravel_index(z.argmax(), z.shape), z.max(), mr.flat[z.argmax()], md.flat[z.argmax()])
    assert (z > 3.0).mean() < 0.05, "%s: %.0f %% of the windows beyond 3 standard errors" % (what, 100 * (z > 3.0).mean())
    for i in range(ref.shape[2]):  # end of training: the whole distribution over runs, not only its mean
        p = sst.ks_2samp(ref[:, -1, i], dev[:, -1, i]).pvalue
        assert p > KS_P_MIN, "%s: agent %d, last window, KS p = %.2g" % (what, i, p)


@pytest.mark.gpu
@pytest.mark.parametrize("name,dtype", [("ensemble_2q", "f32"), ("ensemble_2q", "f64"), ("ensemble_qr", "f32")])
def test_device_ensemble_matches_reference_ensemble(name, dtype):
    import torch
    from th_rl_b200 import trainer
    cfg, window, ref_r, ref_a = _fixture(name)
    E = ref_r.shape[1] * window
    R = 512
    res = trainer.train_many(cfg, R, E, seed=20261018, log_runs=R, dtype=torch.float64 if dtype == "f64" else torch.float32,
                             chunk_epochs=500)
    _compare(ref_r, _windows(res.rewards_log, window), name + " rewards")
    _compare(ref_a, _windows(res.actions_log, window), name + " actions")


def test_oracle_ensemble_matches_reference_ensemble():
    """The same comparison for the oracle's free-running mode (CPU): the Philox restatement the device is bit-compared with
    is itself distributionally faithful to the reference's random streams."""
    from oracle import oracle
    from th_rl_b200 import abi
    cfg, window, ref_r, ref_a = _fixture("ensemble_2q")
    E, R = ref_r.shape[1] * window, 192
    game = oracle.layout(cfg)
    q0, c0, eps0, p0 = oracle.init(game, R, seed=7, dtype=np.float64, eps0=abi.eps0_from_config(cfg))
    out = oracle.scan(game, q0, eps0, p0, E, seed=7, n_log_runs=R, n_threads=oracle.online_cores())
    _compare(ref_r, _windows(out.rewards_log, window), "oracle rewards")
    _compare(ref_a, _windows(out.actions_log, window), "oracle actions")
