# round 2, final build: full GPU suite, DRAM traffic of the conv class stamped with this build's source hash, the bench line, a stress
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2_final_tests.txt; cat gpurun_out/r2_final_tests.txt
B="python bench.py --steps 1 --warmup 3 --skip-cpu --no-graph --inflight 1"
SHA=$(python -c "import bench; print(bench.csrc_sha())")
timeout 600 ncu --kernel-name-base demangled --clock-control none -k 'regex:step::' --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -s 300 -c 320 --csv --log-file gpurun_out/conv_traffic_r2.csv $B > gpurun_out/l2.log 2>&1
python tools/conv_traffic.py gpurun_out/conv_traffic_r2.csv gpurun_out/r2_conv_traffic.json $SHA
python tools/launch_summary.py gpurun_out/conv_traffic_r2.csv full > gpurun_out/r2_launches_summary.txt 2>/dev/null
cp gpurun_out/r2_conv_traffic.json profiles/r2_conv_traffic.json
timeout -s KILL 300 python bench.py --verbose > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r2_bench_n1.json
bash tools/experiments/r2_fused_stress.sh 10 A=1
