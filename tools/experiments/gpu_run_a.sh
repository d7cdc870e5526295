#!/bin/bash
mkdir -p gpurun_out
python -m step_b200.build > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
python tools/conv_bench.py loc_res loc_1088 5b_fused stem_s2d 2>&1 | tail -5
timeout 600 python -m pytest tests -q -m gpu --tb=short -x 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --skip-cpu > gpurun_out/y_bench.log 2>&1; tail -c 1200 gpurun_out/y_bench.log
