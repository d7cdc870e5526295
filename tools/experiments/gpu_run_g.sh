#!/bin/bash
mkdir -p gpurun_out
python -m step_b200.build > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_conv.py -q -m gpu --tb=short -x 2>&1 | tail -5 | tee gpurun_out/g1.log
echo "== tma store"; python tools/conv_bench.py loc_res loc_nores loc_1088 loc_1024 5b_fused 4b_fused 2>&1 | tail -6
echo "== legacy"; STEP_B200_TMAST=0 python tools/conv_bench.py loc_res loc_nores loc_1088 loc_1024 5b_fused 4b_fused 2>&1 | tail -6
timeout 600 python -m pytest tests/test_gpu_pipeline.py -q -m gpu --tb=short -x 2>&1 | tail -3
python bench.py --steps 30 --warmup 5 --skip-cpu > gpurun_out/y_bench.log 2>&1; python - <<'PY'
import json
d=json.loads(open('gpurun_out/y_bench.log').read().strip().splitlines()[-1])
print("clips/s", d["value"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["achieved"], d["roofline"]["ms_per_step_in_kernel"])
PY
