#!/usr/bin/env python
"""bench.py — agent-steps/s of the th_rl training hot path on N B200s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one pass of the hot path (thrl_qtable_scan) over the whole batch: RUNS_PER_GPU independent runs x
2 QTable agents x EPOCHS epochs x 100 env steps on every GPU (BASELINE C2 shape; 8 GPUs x 131,072 runs = the
1,048,576 runs of C3).  Runs are independent, so ranks share nothing on the data path (weak scaling); the only
collective is the NCCL all-reduce of the per-epoch cross-run statistics, inside the timed region.

value  = agent-steps of all ranks / max-over-ranks CUDA-event time, state resident in HBM.
e2e    = the same work through the host-facing API: tables start in pinned HOST memory, are copied in, scanned,
         and tables + counters + statistics are copied back, all inside the timed region.
--impl reference: the CPU arm = the oracle port of the reference path (oracle/thrl_oracle.c, all host threads) on a
         bounded sample of the same workload.  The reference itself is pure Python and cannot travel to the GPU box;
         its measured speed in the build container is quoted in BASELINE.md (2.35e4 agent-steps/s/core).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MAX_STEPS = 100


def _qcfg(n, states, actions, lo, hi, epochs):
    return {
        "agents": [dict(name="QTable", gamma=0.95, actions=actions, states=states, alpha=0.1, eps_end=0.001, epsilon=0.5,
                        eps_step=0.9995, action_range=[lo, hi]) for _ in range(n)],
        "environment": dict(name="NoisyPriceState", noise_prob=0, a=10, b=1, nplayers=n, max_steps=MAX_STEPS),
        "training": dict(print_freq=500, epochs=epochs),
    }


def _c4_hp(R, n):
    """BASELINE C4 sweep grid: alpha x eps_step x gamma = 64 points, repeated over the runs (64 seeds each at R = 4096)."""
    import numpy as np
    grid = [(al, g, st) for al in (.05, .1, .2, .5) for st in (.999, .9995, .9999, .99995) for g in (.35, .8, .95, .99)]
    hp = np.empty((R, n, 4))
    for r in range(R):
        al, g, st = grid[r % 64]
        hp[r, :, 0], hp[r, :, 1], hp[r, :, 2], hp[r, :, 3] = al, g, 0.001, st
    return hp


# name -> workload.  `algo_bytes` = SURVEY 8(d): 8*A + 16 bytes of table traffic per agent-step.  `e2e_chunks`: launches the host
# pipeline cuts a step into (c4: two launches of one full round each -- 4,096 runs are two rounds of the persistent grid, smaller
# launches would leave SMs idle; c5 has only 7 resident waves per step: about one launch per wave).
WORKLOADS = {
    "c2": dict(agents=2, runs_per_gpu=131072, epochs=1000, e2e_chunks=12, config=_qcfg(2, 100, 21, 0.2, 0.4, 1000), algo_bytes=184.0,
               bound="smem", hp=None,
               desc="2-agent QTable iterated Cournot/PD game (example_config hyper-parameters, 101x21 tables, max_steps=100), "
                    "%d runs/GPU x %d epochs per step (C2 shape; 8 GPUs = the 1,048,576 runs of C3)",
               kernel="thrl::qtable_scan_lut2<float, true> (persistent, one launch per step)"),
    "c4": dict(agents=8, runs_per_gpu=4096, epochs=500, e2e_chunks=2, config=_qcfg(8, 1000, 101, 0.05, 0.15, 500), algo_bytes=824.0,
               bound="hbm", hp=_c4_hp,
               desc="hyper-parameter sweep (alpha x eps_step x gamma = 64 points x 64 seeds), 8 QTable agents, 1001x101 "
                    "tables left in HBM, max_steps=100, %d runs/GPU x %d epochs per step (C4 shape)",
               kernel="thrl::qtable_scan_generic<float, false> (persistent, one launch per step)"),
}
def _c5_cfg(epochs):
    a = dict(name="ActorCritic", gamma=0.98, actions=21, states=1, action_range=[0.2, 0.4])  # 1 -> 256 -> {21, 1}, N = 1000
    return {"agents": [dict(a), dict(a)],
            "environment": dict(name="NoisyPriceState", noise_prob=0, a=10, b=1, nplayers=2, max_steps=MAX_STEPS),
            "training": dict(print_freq=500, epochs=epochs)}


# ~1.1e4 flop per act + ~3.4e4 flop per agent-step of amortised update (N = 1000 batch every 10 episodes)
WORKLOADS["c5"] = dict(agents=2, runs_per_gpu=16384, epochs=200, e2e_chunks=8, config=_c5_cfg(200), algo_bytes=4.5e4, bound="tensor", hp=None,
                       desc="2 ActorCritic agents (MLP 1->256->{21,1}, Adam, N=1000 transition batches every 10 episodes), "
                            "%d runs/GPU x %d epochs per step (C5 shape)",
                       kernel="thrl::mlp_scan_pwl (persistent, one launch per step; policy LUT per lattice state + sorted-breakpoint "
                              "gradient sweep, f64 accumulation; no dense contraction is left)")
WL = WORKLOADS["c2"]  # set in main()
CONFIG = WL["config"]
EPOCHS = WL["epochs"]
ALGO_BYTES_PER_AGENT_STEP = WL["algo_bytes"]


def select_workload(name):
    global WL, CONFIG, EPOCHS, ALGO_BYTES_PER_AGENT_STEP
    WL = WORKLOADS[name]
    CONFIG, EPOCHS, ALGO_BYTES_PER_AGENT_STEP = WL["config"], WL["epochs"], WL["algo_bytes"]


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("sm_max_mhz", 1965.0), "measured"
    return 6650.0, 1965.0, "fallback"


def measured_tflops():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("bf16_tflops_sustained", 1400.0), "measured (sustained)"
    return 1400.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True).start()
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_oracle_leg(n_threads, target_seconds=12.0):
    """Times the oracle port on host cores on a bounded sample of the same workload (same config, fewer runs)."""
    import numpy as np
    from oracle import oracle
    from th_rl_b200 import abi
    game = oracle.layout(CONFIG)
    cores = n_threads if n_threads > 0 else oracle.online_cores()
    eps0 = abi.eps0_from_config(CONFIG)
    # calibrate on a small sample, then size the timed sample for ~target_seconds
    R0 = (64 if WL["agents"] == 2 else 1) * cores
    q0, c0, e0, p0, *rest = oracle.init(game, R0, seed=0, dtype=np.float32, eps0=eps0)
    t = time.perf_counter()
    oracle.scan(game, q0, e0, p0, 20, n_threads=cores, n_log_runs=0, stats=True, mlp=rest[0] if rest else None)
    n = WL["agents"]
    rate = R0 * n * 20 * MAX_STEPS / (time.perf_counter() - t)
    R = int(max(cores, min(262144, target_seconds * rate / (n * EPOCHS * MAX_STEPS))))
    q0, c0, e0, p0, *rest = oracle.init(game, R, seed=0, dtype=np.float32, eps0=eps0)
    t = time.perf_counter()
    oracle.scan(game, q0, e0, p0, EPOCHS, n_threads=cores, n_log_runs=0, stats=True, mlp=rest[0] if rest else None)
    dt = time.perf_counter() - t
    return {"value": R * n * EPOCHS * MAX_STEPS / dt, "unit": "agent-steps/s", "cores": cores, "kind": "port",
            "sample": "%d runs x %d epochs x %d steps x %d agents of the bench workload, fp32-storage oracle "
                      "(oracle/thrl_oracle.c), %d pthreads, %.1f s" % (R, EPOCHS, MAX_STEPS, n, cores, dt)}, dt


def _run_stride():
    return sum((a["states"] + 1) * a["actions"] for a in CONFIG["agents"] if a["name"] == "QTable")


def _state_bytes(R):
    from th_rl_b200 import _lib
    g = _lib.game_layout(CONFIG)
    return R * (g.run_stride * 8 + g.mlp_stride * 4)


def base_line(args, n_gpus):
    return {
        "metric": "agent-steps/sec", "unit": "agent-steps/s", "n_gpus": n_gpus, "steps": args.steps,
        "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WL["desc"] % (args.runs_per_gpu, args.epochs), "workload_id": args.workload,
                   "runs_per_gpu": args.runs_per_gpu, "global_runs": args.runs_per_gpu * n_gpus, "epochs_per_step": args.epochs,
                   "max_steps": MAX_STEPS, "agents": WL["agents"], "table_storage": "fp32 (f64 update arithmetic)",
                   "rng": "philox4x32-10", "parallelism": "runs sharded over %d GPU(s), no data-path collective; "
                                                          "NCCL all-reduce of per-epoch statistics" % n_gpus,
                   "l2": "per-GPU state (%.1f GB) is far larger than the 126 MB L2" % (_state_bytes(args.runs_per_gpu) / 1e9)},
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    cb = None
    for i in range(args.warmup + args.steps):
        cb, dt = cpu_oracle_leg(0, target_seconds=max(2.0, min(10.0, 120.0 / (args.warmup + args.steps))))
        if i >= args.warmup:
            vals.append((cb["value"], dt))
    v = statistics.mean(x for x, _ in vals)
    line = base_line(args, args.gpus)
    line.update({"impl": "reference", "value": v, "ms_per_step": 1e3 * statistics.mean(d for _, d in vals),
                 "cpu_baseline": dict(cb, value=v),
                 "e2e": {"value": v, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                 "gpu_launches": 0})
    line["config"]["reference_arm"] = ("oracle port of th_rl/trainer.py:45-70 + agents.py:59-89 + environments.py:25-39 on "
                                       "host cores; the Python reference itself measured 2.35e4 agent-steps/s/core in the "
                                       "build container (BASELINE.md)")
    print(json.dumps(line))


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from th_rl_b200 import _lib, abi, engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    R, E = args.runs_per_gpu, args.epochs
    nag = WL["agents"]
    agent_steps_rank = R * nag * E * MAX_STEPS
    hp = WL["hp"](R, nag) if WL["hp"] else None

    batch = engine.RunBatch(CONFIG, R, device=dev, dtype=torch.float32, seed=0, run_id0=rank * R, hp=hp).init_device()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        out = batch.scan(E, stats=True)
        if world > 1:
            dist.all_reduce(out.stats)  # exact int64 sums: the result does not depend on the sharding
        return out

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    ev[0].record()
    for i in range(args.steps):
        kev[i][0].record()
        out = batch.scan(E, stats=True)
        kev[i][1].record()
        if world > 1:
            dist.all_reduce(out.stats)
        ev[i + 1].record()
    barrier()
    launches = _lib.launch_count() - launches0
    total_ms = ev[0].elapsed_time(ev[-1])
    kern_ms = statistics.mean(a.elapsed_time(b) for a, b in kev)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([total_ms, kern_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, kern_ms = t.tolist()
    value = agent_steps_rank * world * args.steps / (total_ms * 1e-3)

    # ---- e2e: host-resident state in, results back to the host, all inside the timed region
    e2e = measure_e2e(args, torch, np, engine, dev, rank, world, barrier)

    if rank == 0:
        hbm_peak, sm_max_mhz, peak_src = measured_peaks()
        smem_peak_gbs = 128.0 * 148 * sm_max_mhz * 1e6 / 1e9  # 128 B/clk/SM x 148 SMs x max SM clock
        per_gpu_rate = agent_steps_rank / (kern_ms * 1e-3)
        achieved = per_gpu_rate * ALGO_BYTES_PER_AGENT_STEP / 1e9
        line = base_line(args, world)
        line.update({"value": value, "ms_per_step": total_ms / args.steps, "clocks": clocks, "e2e": e2e, "gpu_launches": launches})
        if WL["bound"] == "smem":
            line["roofline"] = {
                "bound": "smem", "achieved": achieved, "peak": smem_peak_gbs, "unit": "GB/s", "frac": achieved / smem_peak_gbs,
                "traffic": None, "kernel": WL["kernel"], "kernel_ms": kern_ms,
                "note": "tables are shared-memory resident, so the bound is SM shared-memory bandwidth / issue rate, not HBM "
                        "(BASELINE.md 5: 184 algorithmic B per agent-step; peak = 128 B/clk/SM x 148 SMs x %.0f MHz from "
                        "MEASURED_PEAKS.json, %s). ncu (profiles/): ~31 warp instructions and ~6 shared-memory wavefronts "
                        "per agent-step, LSU data pipe ~52%% busy, issue slots ~70%% busy; HBM sees only the one-off slab "
                        "load/store and the visit counters" % (sm_max_mhz, peak_src),
                "hbm": {"achieved": per_gpu_rate * (2 * _run_stride() * (4 + 4 + 4) / (nag * 1.0 * E * MAX_STEPS)) / 1e9,
                        "peak": hbm_peak, "unit": "GB/s"},
                # the limiter ncu names: warp-instruction issue.  Instructions per agent-step come from the committed ncu
                # capture of this command (profiles/r1_bench_c2_lut2_ncu_summary.md), the rate is the live one.
                "issue_slots": {"warp_inst_per_agent_step": 30.4, "achieved": per_gpu_rate * 30.4,
                                "peak": 4 * 148 * sm_max_mhz * 1e6, "unit": "warp-inst/s",
                                "frac": per_gpu_rate * 30.4 / (4 * 148 * sm_max_mhz * 1e6)}}
        elif WL["bound"] == "tensor":
            tf_peak, tf_src = measured_tflops()
            tf = per_gpu_rate * ALGO_BYTES_PER_AGENT_STEP / 1e12
            line["roofline"] = {
                "bound": "tensor", "achieved": tf, "peak": tf_peak, "unit": "TFLOP/s", "frac": tf / tf_peak, "traffic": None,
                "kernel": WL["kernel"], "kernel_ms": kern_ms,
                "note": "4.5e4 algorithmic flop per agent-step (BASELINE.md 5: per-step forward + N = 1000 batched update as dense "
                        "GEMMs); peak = bf16 dense from MEASURED_PEAKS.json (%s).  The kernel does NOT execute those flops: the "
                        "network input is a scalar from a finite price lattice, so pi(.|s) is tabulated per state and the exact "
                        "gradient comes from one sorted-breakpoint sweep, O(states*A + H*A) per update instead of O(N*H*A) "
                        "(DESIGN.md 4.5).  `achieved` is therefore the rate the dense formulation would have needed; tensor-pipe "
                        "utilisation is 0 by construction and the real limiter is SM issue rate (profiles/)" % tf_src}
        else:
            line["roofline"] = {
                "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": None, "kernel": WL["kernel"], "kernel_ms": kern_ms,
                "note": "tables (3.2 MB per run) stay in HBM; 824 algorithmic B per agent-step (BASELINE.md 5); peak = measured "
                        "copy bandwidth from MEASURED_PEAKS.json (%s). ncu (profiles/): ~0.87 kB of DRAM traffic and ~123 warp "
                        "instructions per agent-step; the greedy-action cache is carried through the update, so rollouts stop waiting on "
                        "HBM once it is warm; the update pass is a chain of L2 hits at 14 resident warps per SM" % peak_src}
        try:  # DRAM traffic per launch, measured once with ncu --set full on this exact command (profiles/)
            tr = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json"))).get(args.workload)
            if tr and args.runs_per_gpu == WL["runs_per_gpu"]:  # per agent-step as captured x the agent-steps of this launch
                line["roofline"]["traffic"] = tr["traffic_bytes_per_agent_step"] * agent_steps_rank
                line["roofline"]["traffic_source"] = tr["source"]
                line["roofline"]["algorithmic_bytes_per_launch"] = agent_steps_rank * ALGO_BYTES_PER_AGENT_STEP
        except (OSError, ValueError):
            pass
        if args.no_cpu_baseline:
            line["cpu_baseline"] = None
        else:
            line["cpu_baseline"], _ = cpu_oracle_leg(0)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def measure_e2e(args, torch, np, engine, dev, rank, world, barrier):
    """Same metric through the host-facing call: pinned host tables -> device -> scan -> tables, counters, epsilon,
    price and statistics back into pinned host memory.  Runs go through in chunks so copies overlap the kernel."""
    import torch.distributed as dist
    R, E = args.runs_per_gpu, args.epochs
    nag = WL["agents"]
    need = R * _run_stride() * 8 + R * (nag + 1) * 8  # pinned host copy of tables + counters + eps + price
    try:
        avail = int(next(l for l in open("/proc/meminfo") if l.startswith("MemAvailable")).split()[1]) * 1024
    except (OSError, StopIteration):
        avail = 1 << 62
    if need * world * 2 > avail:
        return {"value": None, "unit": "agent-steps/s", "h2d_bytes_per_step": need * world, "d2h_bytes_per_step": need * world,
                "skipped": "host has %.0f GB available, the pinned state needs %.0f GB" % (avail / 1e9, need * world / 1e9)}
    hp = WL["hp"](R, nag) if WL["hp"] else None
    batch = engine.RunBatch(CONFIG, R, device=dev, dtype=torch.float32, seed=1, run_id0=rank * R, hp=hp).init_device()
    torch.cuda.synchronize()
    host = engine.HostState.from_batch(batch)  # pinned host copies of q / counter / eps / price
    steps = max(1, min(args.steps, 3))
    h2d = d2h = 0
    for it in range(1 + steps):  # one warm-up
        if it == 1:
            barrier()
            t0 = time.perf_counter()
        stats, h2d, d2h = engine.scan_from_host(batch, host, E, n_chunks=args.e2e_chunks)
        if world > 1:
            dist.all_reduce(stats)
        stats_h = stats.cpu()  # the step's result (per-epoch cross-run statistics) is read on the host
    barrier()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = t.item()
    return {"value": R * world * nag * E * MAX_STEPS * steps / dt, "unit": "agent-steps/s",
            "h2d_bytes_per_step": int(h2d) * world, "d2h_bytes_per_step": int(d2h + stats_h.numel() * 8) * world,
            "steps": steps, "ms_per_step": 1e3 * dt / steps,
            "api": "th_rl_b200.engine.scan_from_host (pinned host state in, tables+counters+eps+price+stats out)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS), help="c2 = headline (default); c4 = HBM-resident sweep; c5 = MLP (ActorCritic) agents")
    ap.add_argument("--runs-per-gpu", type=int, default=None)
    ap.add_argument("--epochs", type=int, default=None)
    ap.add_argument("--e2e-chunks", type=int, default=None, help="launches the host pipeline cuts a step into (default: per workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    select_workload(args.workload)
    args.runs_per_gpu = args.runs_per_gpu or WL["runs_per_gpu"]
    args.epochs = args.epochs or WL["epochs"]
    args.e2e_chunks = args.e2e_chunks or WL["e2e_chunks"]
    globals()["EPOCHS"] = args.epochs
    if args.impl == "reference":
        return run_reference(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: relaunch under torchrun
        port = 29500 + os.getpid() % 1000
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", str(port)] + sys.argv
        sys.exit(subprocess.call(cmd))
    run_ours(args)


if __name__ == "__main__":
    main()
