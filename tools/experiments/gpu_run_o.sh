#!/bin/bash
mkdir -p gpurun_out
python -m step_b200.build > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests/ -q -m gpu --tb=short -x 2>&1 | tail -8 | tee gpurun_out/o1.log
echo "== halo"; CB_AMODE=4 timeout 120 python tools/conv_bench.py conv2c stem_s2d 2>&1 | tail -2
echo "== im2col"; CB_AMODE=3 timeout 120 python tools/conv_bench.py conv2c 2>&1 | tail -1
timeout 600 python bench.py --steps 30 --warmup 5 --skip-cpu > gpurun_out/y_bench.log 2>&1; python - <<'PY'
import json
d=json.loads(open('gpurun_out/y_bench.log').read().strip().splitlines()[-1])
print("clips/s", d["value"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["achieved"], d["roofline"]["ms_per_step_in_kernel"])
PY
