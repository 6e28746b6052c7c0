"""Attribute an .ncu-rep's per-SASS counters to CUDA source lines: joins `ncu --page source --csv` (SASS order) with
`nvdisasm -g` of the same function in libthrl.so (compiled with -lineinfo).
Usage: python scripts/ncu_by_line.py REP MANGLED_FUNCTION UNITS [top] [instr|stall]   (UNITS = agent-steps in the launch)"""
import csv, io, os, re, subprocess, sys, tempfile

rep, func, units = sys.argv[1], sys.argv[2], float(sys.argv[3])
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
key = 1 if (len(sys.argv) > 5 and sys.argv[5] == "stall") else 0
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "th_rl_b200", "libthrl.so")], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
# walk the function's section: remember the current "//## File ..., line N" annotation for every instruction
start = next(i for i, l in enumerate(sass) if l.startswith(".text." + func + ":"))
lines, cur = [], ("?", 0)
for l in sass[start + 1:]:
    if l.startswith("\t.section") or l.startswith(".text."):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))[2:]
assert len(rows) == len(lines), (len(rows), len(lines))
agg = {}
for (f, ln), r in zip(lines, rows):
    a = agg.setdefault((f, ln), [0, 0])
    a[0] += int(r[5]); a[1] += int(r[2])
tot_i, tot_s = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
text = {}
print("total warp-instr / unit: %.1f" % (tot_i / units))
print("| file:line | warp-instr / unit | %% instr | %% stall samples | source |\n|---|---|---|---|---|")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][key])[:top]:
    if f not in text:
        p = os.path.join(root, "th_rl_b200", "csrc", f)
        text[f] = open(p).read().splitlines() if os.path.exists(p) else []
    srcl = text[f][ln - 1].strip()[:100] if 0 < ln <= len(text[f]) else ""
    print("| %s:%d | %.2f | %.1f | %.1f | `%s` |" % (f, ln, a[0] / units, 100.0 * a[0] / tot_i, 100.0 * a[1] / max(tot_s, 1), srcl))
