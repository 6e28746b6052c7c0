"""Builds th_rl_b200/libthrl.so (the C-ABI library of include/thrl.h) in-tree with nvcc for sm_100a.

    python -m th_rl_b200.build [--force]

--fmad=false: the hot path must round exactly where the reference's numpy / python float operations round
(DESIGN.md "Arithmetic"), so the compiler may not contract a*b+c into an FMA.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libthrl.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = os.environ.get("THRL_EXTRA_NVCC", "").split() + ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "--fmad=false",
         "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall", "-shared", "-Xptxas", "-v"]


def sources():
    return sorted(os.path.join(SRC, f) for f in os.listdir(SRC))


def stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = sources() + [os.path.join(HERE, "..", "include", "thrl.h")]
    return any(os.path.getmtime(p) > t for p in deps)


def build(force=False, verbose=False):
    if not (force or stale()):
        return OUT
    cu = [s for s in sources() if s.endswith(".cu")]
    cmd = [NVCC] + FLAGS + ["-o", OUT] + cu
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log = os.path.join(HERE, "_build", "nvcc.log")
    os.makedirs(os.path.dirname(log), exist_ok=True)
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout)
    if verbose or r.returncode != 0:
        print(r.stdout)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libthrl.so (log: %s)" % log)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
