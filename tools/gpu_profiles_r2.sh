#!/bin/bash
# Round 2: produces everything under gpurun_out/ that profiles/r2_* summarises (one gpurun call, ~12 GPU-minutes).
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --skip-cpu --no-graph --inflight 1"
NCU="timeout 900 ncu --kernel-name-base demangled --clock-control none"
SHA=$(python -c "import bench; print(bench.csrc_sha())")
$NCU -k regex:step:: --metrics gpu__time_duration.sum -s 300 -c 320 --csv --log-file gpurun_out/launches_r2.csv $B > gpurun_out/l.log 2>&1
python tools/launch_summary.py gpurun_out/launches_r2.csv full > gpurun_out/r2_launches_summary.txt
$NCU -k 'regex:step::' --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -s 300 -c 320 --csv --log-file gpurun_out/conv_traffic_r2.csv $B > gpurun_out/l2.log 2>&1
python tools/conv_traffic.py gpurun_out/conv_traffic_r2.csv gpurun_out/r2_conv_traffic.json $SHA
cp gpurun_out/r2_conv_traffic.json profiles/r2_conv_traffic.json   # bench.py reports it only for the build it was captured from
# the pair kernel inside the step: loc_1088 (launch 57), the 3x3x3 Mixed_5 layers, a 1024 -> 256 layer, and the 256 -> 1024 residual layer
$NCU -k 'regex:conv_umma_persist_kernel' --set full --import-source on -s 64 -c 14 -o gpurun_out/prof_r2_persist -f $B > gpurun_out/p2.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_r2_persist.ncu-rep > gpurun_out/r2_ncu_conv_persist_pair.txt 2>&1
$NCU -k 'regex:bottleneck_exit_kernel' --set full --import-source on -s 0 -c 3 -o gpurun_out/prof_r2_exit -f $B > gpurun_out/p4.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_r2_exit.ncu-rep > gpurun_out/r2_ncu_bottleneck_exit.txt 2>&1
$NCU -k 'regex:conv_halo_kernel|clip_to_s2d|detect_' --set full -s 0 -c 3 -o gpurun_out/prof_r2_misc -f $B > gpurun_out/p3.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_r2_misc.ncu-rep > gpurun_out/r2_ncu_stem_s2d_detect.txt 2>&1
timeout -s KILL 300 python tools/conv_bench.py > gpurun_out/r2_conv_layers.txt 2>/dev/null
python tools/microbench.py > gpurun_out/r2_microbench.json 2>/dev/null
timeout -s KILL 400 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
timeout -s KILL 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_n1.json 2>/dev/null
python tools/pool_bench.py > gpurun_out/r2_pool_bench.txt 2>/dev/null
head -30 gpurun_out/r2_launches_summary.txt; cat gpurun_out/r2_conv_traffic.json; cat gpurun_out/r2_ncu_conv_persist_pair.txt | cut -c1-400; cut -c1-900 gpurun_out/r2_bench_n1.json
