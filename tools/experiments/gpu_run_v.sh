#!/bin/bash
mkdir -p gpurun_out
python -m step_b200.build > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_pipeline.py -q -m gpu --tb=short -x 2>&1 | tail -4
timeout 300 python tools/microbench.py > gpurun_out/microbench_v.json 2>gpurun_out/microbench_v.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/microbench_v.json'))
for k,v in d.items():
    if isinstance(v, dict) and k != "peaks": print(k, {a:b for a,b in v.items() if a in ("ms","gb_per_s","frac_of_hbm_peak","clips_per_s","tflops","max_rel_err")})
PY
