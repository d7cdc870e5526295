#!/bin/bash
# round 2: CTA-pair (cta_group::2) persistent kernel vs the round-1 kernels, layer by layer, plus the conv tests with pairs forced
mkdir -p gpurun_out
STEP_B200_PAIR=0 python tools/conv_bench.py > gpurun_out/r2_convbench_pair0.txt 2>&1
STEP_B200_PAIR=1 timeout -s KILL 300 python tools/conv_bench.py > gpurun_out/r2_convbench_pair1.txt 2>&1
timeout -s KILL 300 python tools/conv_bench.py > gpurun_out/r2_convbench_auto.txt 2>&1
STEP_B200_PAIR=1 timeout -s KILL 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_pipeline.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_tests_pair1.txt
paste -d'\n' gpurun_out/r2_convbench_pair0.txt gpurun_out/r2_convbench_pair1.txt | head -80
cat gpurun_out/r2_tests_pair1.txt
