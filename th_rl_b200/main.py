"""main — drop-in for th_rl/main.py:1-26 (click CLI: --runs, --dir; --cdir is the README's spelling of --dir).

    python -m th_rl_b200.main --dir <configs dir> --runs R [--mode batched|reference]

For every *.json in <dir> whose name is not yet present under <dir>/../runs, R runs are trained and written to
<dir>/../runs/<cfg>/<i>/ in the reference's layout.  `--mode batched` (default) plays the R runs as one device scan with
Philox streams; `--mode reference` calls train_one R times with the reference's own host random streams (bit-identical
artefacts under seeded RNGs).
"""
import os

import click

from .trainer import train_many, train_one, _load_config


@click.command()
@click.option("--runs", default=1, help="Runs per config", type=int)
@click.option("--dir", "--cdir", "dir", default=os.path.join(os.getcwd(), "configs"), help="Configs dir", type=str)
@click.option("--mode", default="batched", type=click.Choice(["batched", "reference"]), help="random streams / batching")
@click.option("--seed", default=0, type=int, help="Philox seed (batched mode)")
def main(**params):
    home = os.path.join(os.path.abspath(params["dir"]), "..", "runs")
    if not os.path.exists(home):
        os.mkdir(home)
    for confname in os.listdir(params["dir"]):
        if ".json" in confname:
            cpath = os.path.join(home, confname.replace(".json", ""))
            if confname.replace(".json", "") not in os.listdir(home):
                if not os.path.exists(cpath):
                    os.mkdir(cpath)
                cfgfile = os.path.join(params["dir"], confname)
                if params["mode"] == "reference":
                    for i in range(params["runs"]):
                        train_one(os.path.join(cpath, str(i)), cfgfile)
                else:
                    config = _load_config(cfgfile)
                    train_many(config, params["runs"], seed=params["seed"], export_dir=cpath, export_runs=params["runs"])
                # main.py:22-23: the reference's `else` belongs to its `for`, so this line prints after every config
                print("Skipping {}".format(confname))


if __name__ == "__main__":
    main()
