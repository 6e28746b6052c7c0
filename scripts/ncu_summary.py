"""Summarise an .ncu-rep (one kernel launch, --set full --import-source on) into markdown: headline counters, warp-instruction
budget per agent-step by basic block, stall reasons.  Usage: python scripts/ncu_summary.py REP AGENT_STEPS > profiles/x.md"""
import csv, io, subprocess, sys

rep, agent_steps = sys.argv[1], float(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
get = lambda name: next((vals[i] for i, h in enumerate(hdr) if h == name), "n/a")
unit = lambda name: next((units[i] for i, h in enumerate(hdr) if h == name), "")
print("# ncu summary: %s\n" % get("Kernel Name"))
print("source: `%s` (ncu --set full --clock-control none --import-source on), %.3g agent-steps in the launch\n" % (rep.split("/")[-1], agent_steps))
print("| metric | value |\n|---|---|")
for m in ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
          "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
          "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
          "smsp__warps_eligible.avg.per_cycle_active", "smsp__inst_executed.sum",
          "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
          "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
          "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
          "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_op_red.sum"]:
    print("| %s | %s %s |" % (m, get(m), unit(m)))
try:
    inst = float(get("smsp__inst_executed.sum"))
    wf = float(get("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"))
    dur = float(get("gpu__time_duration.sum")) * (1e-3 if unit("gpu__time_duration.sum") == "ms" else 1e-6 if unit("gpu__time_duration.sum") == "us" else 1e-9)
    print("\n* warp instructions per agent-step: **%.1f**; shared-memory wavefronts per agent-step: %.1f (128 B each -> %.0f GB/s of the 37,225 GB/s smem peak)"
          % (inst / agent_steps, wf / agent_steps, wf * 128 / dur / 1e9))
    print("* agent-steps/s in this (profiled, cold) launch: %.3g" % (agent_steps / dur))
except ValueError:
    pass
srows = list(csv.reader(io.StringIO(src)))
shdr, body = srows[1], srows[2:]
blocks, cur = [], None
for i, r in enumerate(body):
    ex, samp = int(r[5]), int(r[4])
    if cur and cur["ex"] == ex:
        cur["n"] += 1; cur["samp"] += samp; cur["end"] = i
    else:
        cur = dict(ex=ex, n=1, samp=samp, start=i, end=i, first=r[1].strip()); blocks.append(cur)
tsamp = sum(b["samp"] for b in blocks) or 1
blocks.sort(key=lambda b: -b["ex"] * b["n"])
print("\n## hottest basic blocks (SASS)\n\n| SASS rows | instrs | executions / agent-step | warp-instr / agent-step | stall samples |\n|---|---|---|---|---|")
for b in blocks[:12]:
    print("| %d-%d | %d | %.4f | %.2f | %.1f%% |" % (b["start"], b["end"], b["n"], b["ex"] / agent_steps, b["ex"] * b["n"] / agent_steps, 100.0 * b["samp"] / tsamp))
st = [i for i, h in enumerate(shdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = {shdr[i]: 0 for i in st}
for r in body:
    for i in st:
        try: tot[shdr[i]] += int(r[i])
        except ValueError: pass
s = sum(tot.values()) or 1
print("\n## warp stall samples\n")
print(", ".join("%s %.1f%%" % (k, 100.0 * v / s) for k, v in sorted(tot.items(), key=lambda x: -x[1])[:8]))
