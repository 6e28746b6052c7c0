#!/bin/bash
# One ncu --set full capture of the headline kernel on the C2 shape (131072 runs x E epochs); run under gpurun.
E=${1:-40}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:qtable_scan_lut2 --launch-skip 1 -c 1 -f -o gpurun_out/prof_c2 \
  python scripts/quick_time.py 131072 $E > gpurun_out/prof_c2.log 2>&1
