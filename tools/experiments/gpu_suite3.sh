#!/bin/bash
mkdir -p gpurun_out
python -m step_b200.build > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
run() { name=$1; shift; echo "=== $name" ; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n ${TAILN:-6} gpurun_out/$name.log; }
run v1_pipe    python -m pytest tests/test_gpu_pipeline.py -q -m gpu --tb=short
run v2_smoke   python __graft_entry__.py smoke
export STEP_B200_AMODE=im2col
run v3_launches ncu -k regex:step:: --metrics gpu__time_duration.sum --clock-control none -s 310 -c 170 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 3 --skip-cpu --no-graph
run v4_bench_graph   python bench.py --steps 10 --warmup 3 --skip-cpu
run v5_bench_nograph python bench.py --steps 10 --warmup 3 --skip-cpu --no-graph
