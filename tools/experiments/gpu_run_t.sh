#!/bin/bash
mkdir -p gpurun_out
python -m step_b200.build > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_pipeline.py -q -m gpu --tb=short -x 2>&1 | tail -5
CB_AMODE=3 python tools/conv_bench.py 5b_b1b 4e_b1b 4f_b1b 5c_b1b 2>&1 | tail -4
