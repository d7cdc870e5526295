#!/bin/bash
mkdir -p gpurun_out
python -m step_b200.build > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
run() { name=$1; shift; echo "=== $name" ; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n ${TAILN:-8} gpurun_out/$name.log; }
run u1_pipe    python -m pytest tests/test_gpu_pipeline.py -q -m gpu --tb=short
run u2_smoke   python __graft_entry__.py smoke
run u3_bench_box    python bench.py --steps 10 --warmup 3 --skip-cpu
STEP_B200_AMODE=im2col run u4_bench_im2col python bench.py --steps 10 --warmup 3 --skip-cpu
run u5_launches ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 3 --skip-cpu
