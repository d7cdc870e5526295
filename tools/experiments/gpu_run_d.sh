#!/bin/bash
mkdir -p gpurun_out
python -m step_b200.build > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
timeout 300 python -m pytest tests/test_gpu_conv.py -q -m gpu --tb=short -x 2>&1 | tail -4
for c in 1 2; do echo "== CLUSTER=$c"; STEP_B200_CLUSTER=$c timeout 120 python tools/conv_bench.py loc_1088 loc_res loc_1024 5b_fused 2>&1 | tail -4; done
timeout 600 python -m pytest tests -q -m gpu --tb=short -x 2>&1 | tail -3
python bench.py --steps 30 --warmup 5 --skip-cpu --inflight 3 > gpurun_out/y_bench.log 2>&1; python - <<'PY'
import json
d=json.loads(open('gpurun_out/y_bench.log').read().strip().splitlines()[-1])
print("clips/s", d["value"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["achieved"], d["roofline"]["ms_per_step_in_kernel"])
PY
