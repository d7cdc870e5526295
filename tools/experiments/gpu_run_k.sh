#!/bin/bash
mkdir -p gpurun_out
python -m step_b200.build > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_pipeline.py -q -m gpu --tb=short -x 2>&1 | tail -5 | tee gpurun_out/k1.log
LAYERS="loc_res loc_nores loc_1088 loc_1024 5b_fused 4b_fused"
echo "== default"; python tools/conv_bench.py $LAYERS 2>&1 | tail -6
echo "== STW=16"; STEP_B200_STW=16 python tools/conv_bench.py $LAYERS 2>&1 | tail -6
echo "== CLUSTER=2"; STEP_B200_CLUSTER=2 python tools/conv_bench.py $LAYERS 2>&1 | tail -6
echo "== MH=2"; STEP_B200_MH=2 python tools/conv_bench.py $LAYERS 2>&1 | tail -6
python bench.py --steps 30 --warmup 5 --skip-cpu > gpurun_out/y_bench.log 2>&1; python - <<'PY'
import json
d=json.loads(open('gpurun_out/y_bench.log').read().strip().splitlines()[-1])
print("clips/s", d["value"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["achieved"], d["roofline"]["ms_per_step_in_kernel"])
PY
