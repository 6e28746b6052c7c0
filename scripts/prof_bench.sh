#!/bin/bash
# One ncu --set full capture of a workload's dominant kernel inside bench.py; run under gpurun.
# usage: scripts/prof_bench.sh <workload> <kernel regex> [extra bench args]
WL=$1; KR=$2; shift 2
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$KR --launch-skip 2 -c 1 -f -o gpurun_out/prof_$WL \
  python bench.py --workload $WL --steps 1 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/prof_$WL.log 2>&1
