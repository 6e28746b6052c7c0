"""Drop-in behaviour on the GPU: train_one reproduces the reference's artefacts bit for bit under seeded RNGs; the CLI
writes the reference's runs/ layout; train_many's exported runs agree with its device state."""
import json
import os
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_train_one_reproduces_reference_files(golden, tmp_path):
    """Seed python `random` and numpy like the recorded reference run, call train_one: the saved tables, counters and
    log.csv equal what the unmodified reference wrote (th_rl/trainer.py:101-110)."""
    import torch
    from th_rl_b200 import trainer
    cfg = golden["config"]
    cpath = tmp_path / "cfg.json"
    cpath.write_text(json.dumps(cfg))
    seed = int(golden["seed"])
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    out = tmp_path / "run0"
    trainer.train_one(str(out), str(cpath), chunk_epochs=5)
    n = len(cfg["agents"])
    if any(a["name"] != "QTable" for a in cfg["agents"]):
        # Games with a Reinforce agent: its samples come from torch's generator *inside* each step (agents.py:160-163) and
        # cannot be pre-drawn on the host, so the run is not a bit-replay of the reference; the artefacts keep the
        # reference's layout and invariants (trainer.py:101-110, agents.py:215-216).
        E, T = cfg["training"]["epochs"], cfg["environment"]["max_steps"]
        for i, a in enumerate(cfg["agents"]):
            if a["name"] == "QTable":
                t, c = np.load(out / ("%d.npy" % i)), np.load(out / ("%d_counter.npy" % i))
                assert t.shape == golden["q_final_%d" % i].shape and t.dtype == np.float64
                # how many transitions get replayed depends only on min_memory / capacity / max_steps, not on the actions
                assert c.sum() == golden["counter_final_%d" % i].sum()
            else:
                sd = torch.load(out / str(i))
                want = {"Reinforce": ["fc1.weight", "fc1.bias", "fc_pi.weight", "fc_pi.bias"],
                        "ActorCritic": ["fc1.weight", "fc1.bias", "fc_pi.weight", "fc_pi.bias", "fc_v.weight", "fc_v.bias"],
                        "CAC": ["fc1.weight", "fc1.bias", "fc_mu.weight", "fc_mu.bias", "fc_std.weight", "fc_std.bias",
                                "fc_v.weight", "fc_v.bias"]}[a["name"]]
                assert list(sd) == want
                for k, v in sd.items():
                    ref0 = golden["mlp0_%d_%s" % (i, k)]
                    assert v.dtype == torch.float32 and tuple(v.shape) == ref0.shape
                # same torch seed => same initial weights; the saved ones have moved by a few Adam steps of 2e-4
                d = (sd["fc1.weight"].numpy() - golden["mlp0_%d_fc1.weight" % i])
                assert 1e-4 < np.abs(d).max() < 2e-3
        lines = (out / "log.csv").read_text().splitlines()
        assert lines[0] == str(golden["log_header"][0]) and lines[1] == str(golden["log_header"][1]) and len(lines) == 2 + E
        return
    for i in range(n):
        assert np.array_equal(np.load(out / ("%d.npy" % i)), golden["q_final_%d" % i])
        assert np.array_equal(np.load(out / ("%d_counter.npy" % i)), golden["counter_final_%d" % i])
    lines = (out / "log.csv").read_text().splitlines()
    assert lines[0] == str(golden["log_header"][0]) and lines[1] == str(golden["log_header"][1])
    log = np.loadtxt(lines[2:], delimiter=",", ndmin=2)
    assert np.array_equal(log[:, :n], golden["rewards_log"]) and np.array_equal(log[:, n:], golden["actions_log"])
    assert json.loads((out / "config.json").read_text()) == cfg
    assert (out / "config.json").read_text() == json.dumps(cfg, indent=3)


def _small_cfg():
    a = dict(name="QTable", gamma=0.95, actions=21, states=100, alpha=0.1, eps_end=0.001, epsilon=0.5, eps_step=0.9995,
             action_range=[0.2, 0.4])
    return {"agents": [dict(a), dict(a)],
            "environment": dict(name="NoisyPriceState", noise_prob=0, a=10, b=1, nplayers=2, max_steps=100),
            "training": {"epochs": 12, "print_freq": 5}}


def test_cli_writes_reference_layout(tmp_path):
    from click.testing import CliRunner
    from th_rl_b200.main import main
    cdir = tmp_path / "configs"
    cdir.mkdir()
    (cdir / "example_config.json").write_text(json.dumps(_small_cfg()))
    (cdir / "notes.txt").write_text("not a config")
    r = CliRunner().invoke(main, ["--dir", str(cdir), "--runs", "3"])
    assert r.exit_code == 0, r.output
    assert "Skipping example_config.json" in r.output  # the reference's for-else prints this after every config
    runs = tmp_path / "runs" / "example_config"
    assert sorted(os.listdir(runs)) == ["0", "1", "2"]
    for i in range(3):
        assert sorted(os.listdir(runs / str(i))) == ["0.npy", "0_counter.npy", "1.npy", "1_counter.npy", "config.json", "log.csv"]
        t = np.load(runs / str(i) / "0.npy")
        c = np.load(runs / str(i) / "1_counter.npy")
        assert t.shape == (101, 21) and t.dtype == np.float64 and c.dtype == np.float64
        assert c.sum() == 12 * 100  # counter.sum() == epochs * max_steps, as in the shipped sample runs
        lines = (runs / str(i) / "log.csv").read_text().splitlines()
        assert lines[0] == "rewards,rewards,actions,actions" and lines[1] == "0,1,0,1" and len(lines) == 2 + 12
    assert not np.array_equal(np.load(runs / "0" / "0.npy"), np.load(runs / "1" / "0.npy"))
    # second invocation: the config's directory exists -> skipped silently (main.py:16); --cdir is accepted
    r2 = CliRunner().invoke(main, ["--cdir", str(cdir), "--runs", "3"])
    assert r2.exit_code == 0 and r2.output == ""
    # reference mode: one train_one per run, progress lines in the reference's format
    (cdir / "second.json").write_text(json.dumps(_small_cfg()))
    r3 = CliRunner().invoke(main, ["--dir", str(cdir), "--runs", "1", "--mode", "reference"])
    assert r3.exit_code == 0, r3.output
    assert "| episode:  4 | reward:[" in r3.output and "agents:QTable,QTable" in r3.output
    assert sorted(os.listdir(tmp_path / "runs" / "second")) == ["0"]


def test_train_many_matches_single_batch_and_exports(tmp_path):
    import torch
    from th_rl_b200 import abi, engine, trainer
    cfg = _small_cfg()
    res = trainer.train_many(cfg, 40, seed=3, log_runs=40, chunk_epochs=5, export_dir=str(tmp_path / "x"), export_runs=2)
    b = engine.RunBatch(cfg, 40, seed=3).init_device()
    o = b.scan(12, n_log_runs=40, stats=True)
    torch.cuda.synchronize()
    assert torch.equal(res.batch.q, b.q) and torch.equal(res.batch.counter, b.counter)
    assert np.array_equal(res.rewards_log, o.rewards_log.cpu().numpy())
    assert np.array_equal(res.stats, o.stats.cpu().numpy())
    mean_r, sd_r, mean_x, sd_x = res.mean_curves()
    assert np.allclose(mean_r, res.rewards_log.mean(0), atol=1e-8) and np.allclose(sd_r, res.rewards_log.std(0), atol=1e-4)
    assert np.array_equal(np.load(tmp_path / "x" / "1" / "0.npy"), b.tables()[0][1].cpu().numpy().astype(np.float64))
    # learning signal: total reward per step sits between 0 and the cartel level 25 (th_rl/utils.py:91-92)
    assert 0 < mean_r.sum(1).mean() < 25.0


def _shipped_example_config(epochs):
    """th_rl/some_path/configs/example_config.json as shipped (QTable + Reinforce), with training.epochs shortened."""
    return {"agents": [dict(name="QTable", gamma=0.95, actions=21, states=100, alpha=0.1, eps_end=0.001, epsilon=0.5,
                            eps_step=0.9995, action_range=[0.2, 0.4]),
                       dict(name="Reinforce", gamma=0.995, actions=21, states=1, action_range=[0.2, 0.4])],
            "environment": dict(name="NoisyPriceState", noise_prob=0, a=10, b=1, nplayers=2, max_steps=100),
            "training": dict(print_freq=500, epochs=epochs)}


def test_cli_runs_the_shipped_example_config(tmp_path):
    """main.py on the reference's own example config (QTable agent 0, Reinforce agent 1): both CLI modes write the
    reference's artefacts -- `1` is a torch state_dict exactly as th_rl/agents.py:215-216 saves it."""
    import torch
    from click.testing import CliRunner
    from th_rl_b200.main import main
    cdir = tmp_path / "configs"
    cdir.mkdir()
    (cdir / "example_config.json").write_text(json.dumps(_shipped_example_config(30)))
    r = CliRunner().invoke(main, ["--dir", str(cdir), "--runs", "2"])
    assert r.exit_code == 0, r.output
    for i in range(2):
        d = tmp_path / "runs" / "example_config" / str(i)
        assert sorted(os.listdir(d)) == ["0.npy", "0_counter.npy", "1", "config.json", "log.csv"]
        sd = torch.load(d / "1")
        assert {k: tuple(v.shape) for k, v in sd.items()} == {"fc1.weight": (256, 1), "fc1.bias": (256,),
                                                              "fc_pi.weight": (21, 256), "fc_pi.bias": (21,)}
        assert np.load(d / "0_counter.npy").sum() == 30 * 100
        assert len((d / "log.csv").read_text().splitlines()) == 32
    a, b = torch.load(tmp_path / "runs" / "example_config" / "0" / "1"), torch.load(tmp_path / "runs" / "example_config" / "1" / "1")
    assert not torch.equal(a["fc_pi.weight"], b["fc_pi.weight"])  # independent runs
    (cdir / "second.json").write_text(json.dumps(_shipped_example_config(12)))
    r = CliRunner().invoke(main, ["--dir", str(cdir), "--runs", "1", "--mode", "reference"])
    assert r.exit_code == 0, r.output
    assert sorted(os.listdir(tmp_path / "runs" / "second" / "0")) == ["0.npy", "0_counter.npy", "1", "config.json", "log.csv"]


def test_reference_mode_runs_have_independent_policy_samples(tmp_path, monkeypatch):
    """train_one in reference mode takes the Philox key of the MLP agents' action samples from torch's generator: successive
    runs get different keys (so their sampling noise is independent, like the reference's own runs), and reseeding torch
    reproduces a run bit for bit."""
    import torch
    from th_rl_b200 import engine, trainer
    cpath = tmp_path / "cfg.json"
    cpath.write_text(json.dumps(_shipped_example_config(12)))
    seeds, real = [], engine.RunBatch

    def recording(*a, **kw):
        seeds.append(kw.get("seed"))
        return real(*a, **kw)

    monkeypatch.setattr(engine, "RunBatch", recording)
    logs = []
    for k, seed in enumerate((5, None, 5)):
        random.seed(5), np.random.seed(5)
        if seed is not None:
            torch.manual_seed(seed)  # (run 1: torch's generator simply continues)
        out = tmp_path / ("run%d" % k)
        trainer.train_one(str(out), str(cpath))
        logs.append(np.loadtxt((out / "log.csv").read_text().splitlines()[2:], delimiter=",", ndmin=2))
    assert seeds[0] != seeds[1] and seeds[0] == seeds[2] and seeds[0] not in (0, None)
    assert not np.array_equal(logs[0][:, 3], logs[1][:, 3])  # the Reinforce agent's mean actions per epoch
    assert np.array_equal(logs[0], logs[2])
