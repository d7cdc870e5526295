#!/bin/bash
mkdir -p gpurun_out
python -m step_b200.build > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
timeout 300 python -m pytest tests/test_gpu_conv.py -q -m gpu --tb=short -x -k "halo or stem" 2>&1 | tail -15 | tee gpurun_out/m1.log
L="stem_s2d 3b_b2b 3c_b2b 4e_b2b 4f_b2b 5b_b2b conv2c"
echo "== halo"; CB_AMODE=4 timeout 120 python tools/conv_bench.py $L 2>&1 | tail -7
echo "== im2col"; timeout 120 python tools/conv_bench.py $L 2>&1 | tail -7
timeout 600 python bench.py --steps 30 --warmup 5 --skip-cpu > gpurun_out/y_bench.log 2>&1; python - <<'PY'
import json
d=json.loads(open('gpurun_out/y_bench.log').read().strip().splitlines()[-1])
print("clips/s", d["value"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["achieved"], d["roofline"]["ms_per_step_in_kernel"])
PY
