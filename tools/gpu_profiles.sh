#!/bin/bash
# Produces everything under gpurun_out/ that profiles/ summarises.
mkdir -p gpurun_out
python -m step_b200.build > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
B="python bench.py --steps 1 --warmup 3 --skip-cpu --no-graph --inflight 1"
NCU="timeout 600 ncu --kernel-name-base demangled --clock-control none"
$NCU -k regex:step:: --metrics gpu__time_duration.sum -s 300 -c 300 --csv --log-file gpurun_out/launches_r1.csv $B > gpurun_out/l.log 2>&1
$NCU -k 'regex:step::|clip_to_s2d' --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -s 300 -c 300 --csv --log-file gpurun_out/conv_traffic.csv $B > gpurun_out/l2.log 2>&1
python tools/conv_traffic.py gpurun_out/conv_traffic.csv gpurun_out/conv_traffic.json
$NCU -k 'regex:conv_umma_kernel<\(int\)64' --set full --import-source on -s 12 -c 6 -o gpurun_out/prof_conv_v1 -f $B > gpurun_out/p1.log 2>&1
$NCU -k 'regex:conv_umma_persist_kernel' --set full --import-source on -s 30 -c 8 -o gpurun_out/prof_conv_persist -f $B > gpurun_out/p2.log 2>&1
$NCU -k 'regex:conv_halo_kernel' --set full --import-source on -s 3 -c 1 -o gpurun_out/prof_conv_halo -f $B > gpurun_out/p5.log 2>&1
$NCU -k 'regex:maxpool3d|roi_align|clip_to_s2d|linear_mma|mean_mid' --set full -s 40 -c 12 -o gpurun_out/prof_misc -f $B > gpurun_out/p4.log 2>&1
python tools/microbench.py > gpurun_out/microbench.json 2>/dev/null
$NCU -k 'regex:roi_align_fwd_nhwc_f16_packed' --set full --import-source on -c 1 -o gpurun_out/prof_roi_c3 -f python tools/microbench.py > gpurun_out/p3.log 2>&1
python tools/conv_bench.py > gpurun_out/conv_bench.txt 2>/dev/null
mkdir -p profiles; cp gpurun_out/conv_traffic.json profiles/r1_conv_traffic.json
python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 900 gpurun_out/bench_full.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>/dev/null
ls gpurun_out | head -50
