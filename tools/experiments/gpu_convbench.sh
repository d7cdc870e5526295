#!/bin/bash
mkdir -p gpurun_out
python -m step_b200.build > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
for cfg in "CONV=1" "CONV=2 MH=1" "CONV=2 MH=2" "CONV=2 MH=2 STAGES=2" "CONV=2 MH=1 STAGES=3"; do
  echo "=== $cfg"
  env $(for kv in $cfg; do echo STEP_B200_$kv; done) python tools/conv_bench.py 2>&1 | tail -14
done
