#!/usr/bin/env python
"""Stages the UNMODIFIED reference package under oracle/_ref/ so that bench.py can time it on the GPU box's host cores.

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  The reference is pure Python: "building" it is the pip install the bench contract
names, `python -m pip install --no-index --no-build-isolation --no-deps --target oracle/_ref <copy of /root/reference>`
(from a copy under /tmp because /root/reference is read-only and setuptools writes build/ next to setup.py; --no-deps because
plotly / streamlit, which only th_rl/utils.py and dashboard.py import, are not installed and the hot path does not need them).
oracle/_ref/ is git-ignored (no reference source enters the history) but not gpurun-ignored, so it travels to the GPU box like
the built .so files.  Nothing in the product imports it; only bench.py's cpu_baseline / --impl reference legs run it
(oracle/ref_bench.py).  Outcome in this container: "Successfully installed th_rl-0.1".

Usage:  python oracle/stage_reference.py [--reference /root/reference]
"""
import argparse
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
TARGET = os.path.join(HERE, "_ref")


def stage(reference="/root/reference", quiet=True):
    if not os.path.isdir(reference):
        return False
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "reference")
        shutil.copytree(reference, src)
        if os.path.isdir(TARGET):
            shutil.rmtree(TARGET)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links",
               "/opt/wheelhouse", "--target", TARGET, src]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0 or not quiet:
            print(r.stdout)
        if r.returncode != 0:
            raise RuntimeError("pip install of the reference into oracle/_ref failed")
    return True


def staged():
    return os.path.isfile(os.path.join(TARGET, "th_rl", "trainer.py"))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    a = ap.parse_args()
    print("staged" if stage(a.reference, quiet=False) else "no reference at %s" % a.reference)
