#!/bin/bash
# Produces everything under gpurun_out/ that profiles/ summarises.
mkdir -p gpurun_out
python -m step_b200.build > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 600 gpurun_out/bench_full.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>/dev/null
python tools/microbench.py > gpurun_out/microbench.json 2>/dev/null
python tools/conv_bench.py > gpurun_out/conv_bench.txt 2>/dev/null
timeout 600 ncu --kernel-name-base demangled -k regex:step:: --metrics gpu__time_duration.sum --clock-control none -s 300 -c 300 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 3 --skip-cpu --no-graph --inflight 1 > gpurun_out/l.log 2>&1
timeout 600 ncu --kernel-name-base demangled -k 'regex:conv_umma_kernel<\(int\)64' --set full --clock-control none --import-source on -s 60 -c 6 -o gpurun_out/prof_conv_v1 -f python bench.py --steps 1 --warmup 3 --skip-cpu --no-graph --inflight 1 > gpurun_out/p1.log 2>&1
timeout 600 ncu --kernel-name-base demangled -k 'regex:roi_align_fwd_nhwc' --set full --clock-control none --import-source on -c 2 -o gpurun_out/prof_roi_c3 -f python tools/microbench.py > gpurun_out/p3.log 2>&1
ls gpurun_out
