#!/bin/bash
# round 2: per-launch durations of one eager step (ncu, cold-cache, serialised: compare shares, not absolutes)
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --skip-cpu --no-graph --inflight 1"
NCU="timeout 600 ncu --kernel-name-base demangled --clock-control none"
$NCU -k regex:step:: --metrics gpu__time_duration.sum -s 300 -c 320 --csv --log-file gpurun_out/launches_r2.csv $B > gpurun_out/l.log 2>&1
python tools/launch_summary.py gpurun_out/launches_r2.csv full > gpurun_out/r2_launches_summary.txt
head -40 gpurun_out/r2_launches_summary.txt
