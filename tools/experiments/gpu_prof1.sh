#!/bin/bash
mkdir -p gpurun_out
python -m step_b200.build > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
export STEP_B200_AMODE=im2col
timeout 600 ncu --kernel-name-base demangled -k 'regex:conv_umma_kernel<\(int\)64' --set full --clock-control none --import-source on -s 60 -c 8 -o gpurun_out/prof_conv_v1 python bench.py --steps 1 --warmup 3 --skip-cpu --no-graph --inflight 1 > gpurun_out/p1.log 2>&1
timeout 600 ncu --kernel-name-base demangled -k 'regex:conv_umma_persist_kernel<\(int\)64, \(bool\)0' --set full --clock-control none --import-source on -s 150 -c 10 -o gpurun_out/prof_conv_persist python bench.py --steps 1 --warmup 3 --skip-cpu --no-graph --inflight 1 > gpurun_out/p2.log 2>&1
ls -la gpurun_out/*.ncu-rep
