"""SURVEY 8(f) rank 2 -- packed on-disk state and cross-run quantile statistics (th_rl/utils.py:132-145).
CPU: file format round trip, the numpy restatement of thrl_curve_hist against pandas' ewm + numpy.quantile.
GPU: thrl_curve_hist == the restatement bit for bit; train_many's histogram / quantile curves; checkpoint -> resume equals an
uninterrupted run; runs exported in the legacy layout load with the reference's own utils.load_experiment."""
import json
import os
import sys
import types

import numpy as np
import pytest

from oracle import oracle
from th_rl_b200 import checkpoint, stats

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cfg2(T=20, noise=0.0, **agent):
    a = dict(name="QTable", gamma=0.95, actions=21, states=100, alpha=0.1, eps_end=0.001, epsilon=0.5, eps_step=0.9995,
             action_range=[0.2, 0.4], min_memory=T)
    a.update(agent)
    return {"agents": [dict(a), dict(a)],
            "environment": dict(name="NoisyPriceState", noise_prob=noise, a=10, b=1, nplayers=2, max_steps=T),
            "training": dict(print_freq=500, epochs=6)}


# ------------------------------------------------------------------------------------------------------------------ CPU
def test_pack_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    arrays = {"q": rng.normal(size=(7, 13)).astype(np.float32), "counter": rng.integers(0, 2 ** 32, (7, 13), dtype=np.uint32),
              "eps": rng.random((7, 2)), "price": rng.random(7), "empty": np.zeros((0, 4), np.float32),
              "ring": rng.integers(0, 255, (7, 33), dtype=np.uint8)}
    hdr = dict(config=_cfg2(), n_runs=7, epoch=12, extra={"k": [1, 2.5, "x"]})
    p = str(tmp_path / "state.thrl")
    size = checkpoint.write_pack(p, hdr, arrays)
    assert os.path.getsize(p) >= size - checkpoint.ALIGN and not os.path.exists(p + ".tmp")
    for mm in (True, False):
        h, a = checkpoint.read_pack(p, mmap=mm)
        assert h["config"] == hdr["config"] and h["epoch"] == 12 and h["extra"] == hdr["extra"]
        assert set(a) == set(arrays)
        for k in arrays:
            assert a[k].dtype == arrays[k].dtype and a[k].shape == arrays[k].shape and np.array_equal(np.asarray(a[k]), arrays[k])
    for m in h["arrays"]:
        assert m["offset"] % checkpoint.ALIGN == 0
    open(p, "r+b").write(b"NOTAPACK")
    with pytest.raises(ValueError):
        checkpoint.read_pack(p)


def test_ewm_denominators_carry():
    d = stats.ewm_decay(1000.0)
    assert abs(d - 0.5 ** (1 / 1000.0)) < 1e-15
    whole = stats.ewm_denominators(d, 0, 50)
    a = stats.ewm_denominators(d, 0, 20)
    b = stats.ewm_denominators(d, 20, 30, den_before=a[-1])
    c = stats.ewm_denominators(d, 20, 30)  # recomputed from epoch 0
    assert np.array_equal(whole, np.concatenate([a, b])) and np.array_equal(b, c)


def test_curve_hist_restatement_matches_pandas_quantiles():
    """oracle.curve_hist + stats.quantiles_from_hist against the reference's own recipe (utils.py:136-145): pandas ewm(halflife)
    of every agent's column, summed, numpy / pandas quantiles over the runs -- equal to within one bin."""
    import pandas
    rng = np.random.default_rng(3)
    R, E, n, hl, bins = 301, 80, 2, 7.0, 2048
    logs = 11.0 + rng.normal(0, 1.5, (R, E, n)).cumsum(axis=1) * 0.05 + rng.normal(0, 0.3, (R, E, n))
    lo, hi = 0.0, 30.0
    d = stats.ewm_decay(hl)
    # two chunks: the numerator and the denominator are carried
    h1, num = oracle.curve_hist(logs[:, :30], d, stats.ewm_denominators(d, 0, 30), np.zeros(R), lo, hi, bins)
    h2, num = oracle.curve_hist(logs[:, 30:], d, stats.ewm_denominators(d, 30, E - 30), num, lo, hi, bins)
    hist = np.concatenate([h1, h2])
    h_all, _ = oracle.curve_hist(logs, d, stats.ewm_denominators(d, 0, E), np.zeros(R), lo, hi, bins)
    assert np.array_equal(hist, h_all) and np.all(hist.sum(1) == R)
    smooth = np.stack([pandas.DataFrame(logs[r]).ewm(halflife=hl).mean().sum(axis=1).to_numpy() for r in range(R)], axis=1)  # [E, R]
    want = np.quantile(smooth, [0.5, 0.75, 0.25], axis=1).T
    got = stats.quantiles_from_hist(hist, (0.5, 0.75, 0.25), lo, hi)
    assert np.abs(got - want).max() <= 1.01 * (hi - lo) / bins


def test_quantiles_from_hist_small_cases():
    h = np.array([[0, 3, 0, 1], [4, 0, 0, 0], [0, 0, 0, 0]])
    q = stats.quantiles_from_hist(h, (0.0, 0.5, 1.0), 0.0, 4.0)
    assert np.allclose(q[0], [1.5, 1.5, 3.5]) and np.allclose(q[1], [0.5, 0.5, 0.5]) and np.all(np.isnan(q[2]))


# ------------------------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 2, 3])
def test_curve_hist_kernel_equals_restatement(n):
    import torch
    rng = np.random.default_rng(10 + n)
    R, E, bins = 1000, 37, 64
    logs = rng.normal(12, 6, (R, E, n)) / n
    logs[5, 3, 0] = np.nan       # lands in the last bin and stays there (the numerator is NaN from then on)
    logs[6, :, :] = -50.0        # below the range: bin 0
    logs[7, :, :] = 1e6          # above: last bin
    cfg = _cfg2()
    ch = stats.CurveHistogram(cfg, R, "cuda:0", halflife=5.0, bins=bins, value_range=(0.0, 30.0))
    hist = torch.zeros((E, bins), dtype=torch.int64, device="cuda:0")
    dev_logs = torch.from_numpy(logs).cuda()
    for b, e in ((0, 10), (10, 11), (11, E)):  # chunked updates carry the EWM state
        ch.update(dev_logs[:, b:e].contiguous(), hist[b:e])
    torch.cuda.synchronize()
    want, num = oracle.curve_hist(logs, ch.decay, stats.ewm_denominators(ch.decay, 0, E), np.zeros(R), 0.0, 30.0, bins)
    assert np.array_equal(hist.cpu().numpy(), want)
    assert np.array_equal(ch.num.cpu().numpy(), num, equal_nan=True)
    # shards add up
    h2 = torch.zeros_like(hist)
    for lo_, hi_ in ((0, 333), (333, R)):
        c2 = stats.CurveHistogram(cfg, hi_ - lo_, "cuda:0", halflife=5.0, bins=bins, value_range=(0.0, 30.0))
        c2.update(dev_logs[lo_:hi_].contiguous(), h2)
    assert torch.equal(h2, hist)


@pytest.mark.gpu
def test_train_many_quantile_curves():
    from th_rl_b200 import trainer
    cfg = _cfg2()
    R, E = 500, 12
    res = trainer.train_many(cfg, R, E, seed=3, log_runs=R, quantile_bins=512, halflife=4.0, chunk_epochs=5)
    d = stats.ewm_decay(4.0)
    lo, hi = stats.default_range(cfg)
    assert (lo, hi) == (0.0, 30.0)
    want, _ = oracle.curve_hist(res.rewards_log, d, stats.ewm_denominators(d, 0, E), np.zeros(R), lo, hi, 512)
    assert np.array_equal(res.curve_hist, want)
    import pandas
    smooth = np.stack([pandas.DataFrame(res.rewards_log[r]).ewm(halflife=4.0).mean().sum(axis=1).to_numpy() for r in range(R)], axis=1)
    q = res.quantile_curves()
    assert np.abs(q - np.quantile(smooth, [0.5, 0.75, 0.25], axis=1).T).max() <= 1.01 * (hi - lo) / 512


def _mixed_cfg():
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "mixed_arq_seed12.npz")))
    return json.loads(str(g["config"]))


def _ring_cfg():
    c = _cfg2(T=12)
    c["agents"][0].update(min_memory=30, capacity=40)   # batches span episodes: pending transitions live in the ring
    c["agents"][1].update(min_memory=12, capacity=9)
    return c


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["qtable", "mlp", "ring", "c4"])
def test_checkpoint_resume_equals_uninterrupted_run(case, tmp_path):
    import torch
    from th_rl_b200 import engine, trainer
    cfg = {"qtable": _cfg2, "mlp": _mixed_cfg, "ring": _ring_cfg,
           "c4": lambda: {"agents": [dict(name="QTable", gamma=0.95, actions=101, states=1000, alpha=0.1, eps_end=0.001, epsilon=0.5,
                                           eps_step=0.9995, action_range=[0.05, 0.15]) for _ in range(8)],
                          "environment": dict(name="NoisyPriceState", noise_prob=0, a=10, b=1, nplayers=8, max_steps=30),
                          "training": dict(print_freq=500, epochs=4)}}[case]()
    R, E1, E2 = (6 if case == "c4" else 40), 7, 5
    hp = None
    if case == "qtable":
        hp = np.tile(np.array([0.2, 0.9, 0.01, 0.999]), (R, 2, 1)) * np.linspace(0.5, 1.0, R)[:, None, None]
    kw = dict(seed=11, hp=hp, quantile_bins=128, halflife=3.0)
    whole = trainer.train_many(cfg, R, E1 + E2, **kw)
    p = str(tmp_path / "shard{rank}.thrl")
    first = trainer.train_many(cfg, R, E1, checkpoint_to=p, **kw)
    assert os.path.exists(p.format(rank=0))
    del first
    second = trainer.train_many(cfg, R, E2, resume_from=p, **kw)
    a, b = whole.batch, second.batch
    assert b.epoch == E1 + E2
    assert torch.equal(a.q, b.q) and torch.equal(a.counter, b.counter) and torch.equal(a.eps, b.eps) and torch.equal(a.price, b.price)
    if a.mlp is not None:
        assert torch.equal(a.mlp.view(torch.int32), b.mlp.view(torch.int32))
    if a.ring is not None:
        assert torch.equal(a.ring, b.ring)
    assert np.array_equal(whole.stats[E1:], second.stats)
    assert np.array_equal(whole.curve_hist[E1:], second.curve_hist)
    hdr, arrays = checkpoint.read_pack(p.format(rank=0))
    assert hdr["epoch"] == E1 and hdr["n_runs"] == R and hdr["config"] == cfg and "x_curve_num" in arrays
    with pytest.raises(ValueError):
        trainer.train_many(cfg, R + 1, E2, resume_from=p, **kw)  # another sharding


@pytest.mark.gpu
def test_exported_runs_load_with_the_reference_utils(tmp_path):
    """th_rl/utils.py:12-24 load_experiment (the UNMODIFIED reference, staged under oracle/_ref; plotly, which this image lacks
    and only the plotting functions use, is stubbed) reads a run that train_many exported: same tables, same smoothed curves."""
    ref = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.isfile(os.path.join(ref, "th_rl", "utils.py")):
        pytest.skip("oracle/_ref is not staged (python oracle/stage_reference.py)")
    import pandas
    from th_rl_b200 import trainer
    cfg = _cfg2(T=100, min_memory=100)
    cfg["training"]["epochs"] = 8
    res = trainer.train_many(cfg, 3, 8, seed=5, log_runs=3, export_dir=str(tmp_path / "runs" / "cfg"), export_runs=3)
    for name in ("plotly", "plotly.express", "plotly.graph_objects", "plotly.subplots"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["plotly.subplots"].make_subplots = lambda *a, **k: None
    sys.path.insert(0, ref)
    try:
        for k in [k for k in sys.modules if k == "th_rl" or k.startswith("th_rl.")]:
            del sys.modules[k]
        from th_rl import utils as ref_utils
        for r in range(3):
            config, agents, environment, actions, rewards = ref_utils.load_experiment(str(tmp_path / "runs" / "cfg" / str(r)))
            assert config == cfg and len(agents) == 2
            for i, a in enumerate(agents):
                assert np.array_equal(a.table, res.batch.tables()[i][r].cpu().numpy().astype(np.float64))
                assert np.array_equal(a.counter, res.batch.counters()[i][r].cpu().numpy().view(np.uint32).astype(np.float64))
            # pandas.read_csv takes the second header row of log.csv ("0,1,0,1", trainer.py:105-110) for a data row
            rows = np.concatenate([[[0.0, 1.0]], res.rewards_log[r]])
            want = pandas.DataFrame(rows).ewm(halflife=1000).mean().to_numpy()
            assert rewards.shape == want.shape and np.allclose(rewards.to_numpy(), want, rtol=0, atol=1e-12)
    finally:
        sys.path.remove(ref)
        for k in [k for k in sys.modules if k == "th_rl" or k.startswith("th_rl.")]:
            del sys.modules[k]
