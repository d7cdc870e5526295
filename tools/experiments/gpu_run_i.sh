#!/bin/bash
mkdir -p gpurun_out
python -m step_b200.build > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
STEP_B200_EPIW=16 timeout 600 python -m pytest tests/test_gpu_conv.py -q -m gpu --tb=short -x 2>&1 | tail -3
STEP_B200_EPIW=16 STEP_B200_SLABBUFS=1 timeout 600 python -m pytest tests/test_gpu_conv.py -q -m gpu --tb=short -x 2>&1 | tail -3
LAYERS="loc_res loc_nores loc_1088 loc_1024 5b_fused 4b_fused"
for w in 8 16; do for b in 1 2; do
echo "== EPIW=$w SLABBUFS=$b"; STEP_B200_EPIW=$w STEP_B200_SLABBUFS=$b python tools/conv_bench.py $LAYERS 2>&1 | tail -6
done; done
for w in 8 16; do
STEP_B200_EPIW=$w python bench.py --steps 30 --warmup 5 --skip-cpu > gpurun_out/y_bench_$w.log 2>&1; python - $w <<'PY'
import json, sys
d=json.loads(open('gpurun_out/y_bench_%s.log' % sys.argv[1]).read().strip().splitlines()[-1])
print("EPIW", sys.argv[1], "clips/s", d["value"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["achieved"], d["roofline"]["ms_per_step_in_kernel"])
PY
done
