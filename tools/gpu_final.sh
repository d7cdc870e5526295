#!/bin/bash
mkdir -p gpurun_out
python -m step_b200.build > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
timeout 1200 python -m pytest tests/ -q -m gpu --tb=short 2>&1 | tail -6 | tee gpurun_out/final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
B="python bench.py --steps 1 --warmup 3 --skip-cpu --no-graph --inflight 1"
timeout 600 ncu --kernel-name-base demangled --clock-control none -k 'regex:conv_halo_kernel<\(int\)32, \(int\)4' --set full --import-source on -s 3 -c 1 -o gpurun_out/prof_conv_halo_stem -f $B > gpurun_out/p5.log 2>&1
ls -la gpurun_out/prof_conv_halo_stem.ncu-rep
