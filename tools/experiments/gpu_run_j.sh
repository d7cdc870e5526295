#!/bin/bash
mkdir -p gpurun_out
python -m step_b200.build > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
for d in 0 7 39 71 135 231 8; do
echo "== DBG=$d"; STEP_B200_DBG=$d python tools/conv_bench.py loc_nores loc_1088 2>&1 | tail -2
done
