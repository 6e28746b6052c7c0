"""Warp instructions and stall samples of an mlp_scan_pwc capture by code region (markers are looked up in the source, so the
table follows edits).  Usage: python scripts/ncu_regions_pwc.py REP MANGLED UNITS"""
import re, subprocess, sys, os
rep, func, units = sys.argv[1], sys.argv[2], sys.argv[3]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run([sys.executable, os.path.join(root, "scripts", "ncu_by_line.py"), rep, func, units, "5000"], capture_output=True, text=True).stdout
lines = open(os.path.join(root, "th_rl_b200", "csrc", "thrl_scan_pwc.cuh")).read().splitlines()
def find(t): return next(i + 1 for i, l in enumerate(lines) if t in l)
marks = [("helpers", 1), ("build (thresholds, ranking, table)", find("__device__ inline void pwc_build")), ("sample (general loop)", find("__device__ __forceinline__ int pwc_sample")),
         ("update: event sort", find("__device__ inline int pwc_sort_events")), ("update: nan", find("__device__ inline void pwc_nan_update")),
         ("update: coefficients", find("__device__ inline void pwc_train(")), ("update: sweep", find("// ---- 3. the sweep")),
         ("update: units", find("// ---- 4. lane = rank q")), ("kernel prologue", find("__global__ void __launch_bounds__(512, 1) mlp_scan_pwc")),
         ("draws", find("// ---- per-episode draws")), ("episode", find("// ---- the episode (trainer.py:50-67)")),
         ("after episode", find("// ---- train_net for every agent in order"))]
tot = {}
for l in out.splitlines():
    m = re.match(r"\| (\S+):(\d+) \| ([\d.]+) \| ([\d.]+) \| ([\d.]+)", l)
    if not m:
        continue
    f, ln, v, ps = m.group(1), int(m.group(2)), float(m.group(3)), float(m.group(5))
    key = f
    if f == "thrl_scan_pwc.cuh":
        for k, a in marks:
            if ln >= a:
                key = k
    t = tot.setdefault(key, [0, 0]); t[0] += v; t[1] += ps
print("| region | warp instructions / agent-step | stall samples |\n|---|---|---|")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    if v[0] >= 0.05:
        print("| %s | %.1f | %.1f %% |" % (k, v[0], v[1]))
