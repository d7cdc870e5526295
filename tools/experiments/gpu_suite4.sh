#!/bin/bash
mkdir -p gpurun_out
python -m step_b200.build > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
run() { name=$1; shift; echo "=== $name" ; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n ${TAILN:-6} gpurun_out/$name.log; }
export STEP_B200_AMODE=${AMODE:-im2col}
run x1_tests   python -m pytest tests -q -m gpu --tb=short -x
run x2_bench   python bench.py --steps 10 --warmup 3 --skip-cpu
run x2b_bench_1flight python bench.py --steps 10 --warmup 3 --skip-cpu --inflight 1
run x3_launches ncu --kernel-name-base demangled -k regex:step:: --metrics gpu__time_duration.sum --clock-control none -s 300 -c 300 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 3 --skip-cpu --no-graph
python tools/launch_summary.py gpurun_out/launches_r1.csv | head -16
