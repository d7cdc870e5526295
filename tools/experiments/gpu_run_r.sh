#!/bin/bash
mkdir -p gpurun_out
python -m step_b200.build > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests/ -q -m gpu --tb=short -x 2>&1 | tail -5 | tee gpurun_out/r1.log
LAYERS="loc_res loc_nores loc_1088 loc_1024 5b_fused 4b_fused"
echo "== STW=32"; python tools/conv_bench.py $LAYERS 2>&1 | tail -6
echo "== STW=16"; STEP_B200_STW=16 python tools/conv_bench.py $LAYERS 2>&1 | tail -6
timeout 300 python tools/microbench.py > gpurun_out/microbench_r.json 2>gpurun_out/microbench_r.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/microbench_r.json'))
for k,v in d.items():
    if isinstance(v, dict): print(k, {a:b for a,b in v.items() if a in ("ms","gb_per_s","frac_of_hbm_peak","clips_per_s","tflops")})
PY
timeout 600 python bench.py --steps 30 --warmup 5 --skip-cpu > gpurun_out/y_bench.log 2>&1; python - <<'PY'
import json
d=json.loads(open('gpurun_out/y_bench.log').read().strip().splitlines()[-1])
print("clips/s", d["value"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["achieved"], d["roofline"]["ms_per_step_in_kernel"])
PY
