#!/bin/bash
mkdir -p gpurun_out
python -m step_b200.build > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_pipeline.py -q -m gpu --tb=short -x 2>&1 | tail -5 | tee gpurun_out/s1.log
L="3b_b1b 4c_b1b 4e_b1b 4f_b1b 5b_b1b 5c_b1b 5b_b2b 5c_b2b 3c_b2b loc_1088 5b_fused 4b_fused"
echo "== default BK policy"; CB_AMODE=3 python tools/conv_bench.py $L 2>&1 | tail -12
echo "== BKPOL=1"; STEP_B200_BKPOL=1 CB_AMODE=3 python tools/conv_bench.py $L 2>&1 | tail -12
timeout 600 python bench.py --steps 30 --warmup 5 --skip-cpu > gpurun_out/y_bench.log 2>&1; python - <<'PY'
import json
d=json.loads(open('gpurun_out/y_bench.log').read().strip().splitlines()[-1])
print("clips/s", d["value"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["achieved"], d["roofline"]["ms_per_step_in_kernel"])
PY
STEP_B200_BKPOL=1 timeout 600 python bench.py --steps 30 --warmup 5 --skip-cpu > gpurun_out/y_bench1.log 2>&1; python - <<'PY'
import json
d=json.loads(open('gpurun_out/y_bench1.log').read().strip().splitlines()[-1])
print("BKPOL=1 clips/s", d["value"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["achieved"], d["roofline"]["ms_per_step_in_kernel"])
PY
