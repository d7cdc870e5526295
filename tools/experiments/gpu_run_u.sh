#!/bin/bash
mkdir -p gpurun_out
python -m step_b200.build > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
L="4c_b1b 4e_b1b 4f_b1b 4e_b2b 4f_b2b"
echo "== default"; CB_AMODE=3 python tools/conv_bench.py $L 2>&1 | tail -5
echo "== NSPLIT=1"; STEP_B200_NSPLIT=1 CB_AMODE=3 python tools/conv_bench.py $L 2>&1 | tail -5
b() { python - "$1" <<'PY'
import json, sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "clips/s", d["value"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["achieved"], d["roofline"]["ms_per_step_in_kernel"])
PY
}
timeout 600 python bench.py --steps 30 --warmup 5 --skip-cpu > gpurun_out/yu0.log 2>&1; b gpurun_out/yu0.log
STEP_B200_NSPLIT=1 timeout 600 python bench.py --steps 30 --warmup 5 --skip-cpu > gpurun_out/yu1.log 2>&1; b gpurun_out/yu1.log
