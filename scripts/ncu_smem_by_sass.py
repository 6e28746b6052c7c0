"""Per-SASS-instruction shared-memory wavefronts (actual vs ideal) and short-scoreboard stall samples of an .ncu-rep.
Usage: python scripts/ncu_smem_by_sass.py REP [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
out = []
tot_w = tot_i = tot_s = 0
for k, r in enumerate(rows[2:]):
    if len(r) < len(hdr): continue
    w, wi = float(r[ix["L1 Wavefronts Shared"]] or 0), float(r[ix["L1 Wavefronts Shared Ideal"]] or 0)
    ssb = float(r[ix["stall_short_sb"]] or 0)
    smp = float(r[ix["# Samples"]] or 0)
    tot_w += w; tot_i += wi; tot_s += smp
    out.append((w - wi, w, wi, ssb, smp, k, r[ix["Source"]].strip()))
print("shared wavefronts %.3g, ideal %.3g, samples %d" % (tot_w, tot_i, tot_s))
print("-- by excess wavefronts")
for e, w, wi, ssb, smp, k, s in sorted(out, reverse=True)[:top]:
    print("%5d excess %.3g (%.1f%% of all wavefronts) actual %.3g ideal %.3g  samples %.1f%%  %s" % (k, e, 100 * e / tot_w, w, wi, 100 * smp / tot_s, s))
print("-- by samples")
for e, w, wi, ssb, smp, k, s in sorted(out, key=lambda x: -x[4])[:top]:
    print("%5d samples %.1f%% (short_sb %d)  %s" % (k, 100 * smp / tot_s, ssb, s))
