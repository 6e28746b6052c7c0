"""Trimmed SASS listing of the hottest basic blocks of an .ncu-rep (one launch, --import-source on): for each of the top K blocks
(by executed warp instructions) every SASS row with its execution count and stall samples.
Usage: python scripts/ncu_hot_sass.py REP UNITS [K] >> profiles/x.md   (UNITS = agent-steps in the launch)"""
import csv, io, subprocess, sys

rep, units = sys.argv[1], float(sys.argv[2])
K = int(sys.argv[3]) if len(sys.argv) > 3 else 3
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
body = rows[2:]
blocks, cur = [], None
for i, r in enumerate(body):
    ex = int(r[5])
    if cur and cur["ex"] == ex:
        cur["end"] = i
    else:
        cur = dict(ex=ex, start=i, end=i); blocks.append(cur)
blocks.sort(key=lambda b: -b["ex"] * (b["end"] - b["start"] + 1))
print("\n## SASS of the %d hottest basic blocks\n" % K)
for b in blocks[:K]:
    n = b["end"] - b["start"] + 1
    print("### SASS rows %d-%d: %d instructions x %.4f executions per agent-step = %.1f warp instructions per agent-step\n\n```" % (
        b["start"], b["end"], n, b["ex"] / units, b["ex"] * n / units))
    for r in body[b["start"]:b["end"] + 1]:
        print("%-90s  # stall samples %s" % (r[1].strip()[:90], r[4]))
    print("```\n")
