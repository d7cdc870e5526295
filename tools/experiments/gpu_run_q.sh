#!/bin/bash
mkdir -p gpurun_out
python -m step_b200.build > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_pipeline.py -q -m gpu --tb=short -x 2>&1 | tail -5 | tee gpurun_out/q1.log
b() { python - "$1" <<'PY'
import json, sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "clips/s", d["value"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["achieved"], d["roofline"]["ms_per_step_in_kernel"])
PY
}
timeout 600 python bench.py --steps 30 --warmup 5 --skip-cpu > gpurun_out/yq_march.log 2>&1; b gpurun_out/yq_march.log
STEP_B200_POOLMARCH=0 timeout 600 python bench.py --steps 30 --warmup 5 --skip-cpu > gpurun_out/yq_nomarch.log 2>&1; b gpurun_out/yq_nomarch.log
timeout 600 python bench.py --steps 30 --warmup 5 --skip-cpu --inflight 4 > gpurun_out/yq_if4.log 2>&1; b gpurun_out/yq_if4.log
timeout 600 python bench.py --steps 30 --warmup 5 --skip-cpu --inflight 2 > gpurun_out/yq_if2.log 2>&1; b gpurun_out/yq_if2.log
