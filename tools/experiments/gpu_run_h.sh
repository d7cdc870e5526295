#!/bin/bash
mkdir -p gpurun_out
python -m step_b200.build > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_conv.py -q -m gpu --tb=short 2>&1 | tail -40 > gpurun_out/h1.log
STEP_B200_DEBUG_SYNC=1 CUDA_LAUNCH_BLOCKING=1 timeout 600 python -m pytest tests/test_gpu_pipeline.py -q -m gpu --tb=short -x -k "pipe_c1" 2>&1 | tail -60 > gpurun_out/h2.log
tail -5 gpurun_out/h1.log; grep -n "conv fault\|Error\|error\|assert" gpurun_out/h2.log | head
