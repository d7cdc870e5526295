#!/bin/bash
mkdir -p gpurun_out
python -m step_b200.build > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
B="python bench.py --steps 1 --warmup 3 --skip-cpu --no-graph --inflight 1"
timeout 600 ncu --kernel-name-base demangled --clock-control none -k 'regex:conv_halo_kernel<\(int\)32, \(int\)4' --set full --import-source on -s 6 -c 1 -o gpurun_out/prof_conv_halo_stem -f $B > gpurun_out/p5.log 2>&1
ls -la gpurun_out/prof_conv_halo_stem.ncu-rep
