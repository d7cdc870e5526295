#!/bin/bash
mkdir -p gpurun_out
python -m step_b200.build > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
L="3b_b2b 3c_b2b 4e_b2b 4f_b2b 5b_b2b"
echo "== halo"; CB_AMODE=4 timeout 120 python tools/conv_bench.py $L 2>&1 | tail -5
timeout 600 ncu --kernel-name-base demangled -k regex:step:: --metrics gpu__time_duration.sum --clock-control none -s 300 -c 300 --csv --log-file gpurun_out/launches_n.csv python bench.py --steps 1 --warmup 3 --skip-cpu --no-graph --inflight 1 > gpurun_out/n.log 2>&1
python tools/launch_summary.py gpurun_out/launches_n.csv x 2>/dev/null | head -190
