#!/bin/bash
mkdir -p gpurun_out
python -m step_b200.build > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests/ -q -m gpu --tb=short -x 2>&1 | tail -8 | tee gpurun_out/l1.log
python bench.py --steps 30 --warmup 5 --skip-cpu > gpurun_out/y_bench.log 2>&1; python - <<'PY'
import json
d=json.loads(open('gpurun_out/y_bench.log').read().strip().splitlines()[-1])
print("clips/s", d["value"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["achieved"], d["roofline"]["ms_per_step_in_kernel"])
PY
STEP_B200_POOL333=0 python bench.py --steps 30 --warmup 5 --skip-cpu > gpurun_out/y_bench0.log 2>&1; python - <<'PY'
import json
d=json.loads(open('gpurun_out/y_bench0.log').read().strip().splitlines()[-1])
print("old pool: clips/s", d["value"], "e2e", d["e2e"]["value"])
PY
ncu --kernel-name-base demangled --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_l.csv python bench.py --steps 1 --warmup 3 --skip-cpu --no-graph --inflight 1 > gpurun_out/l2.log 2>&1
python tools/launch_summary.py gpurun_out/launches_l.csv 2>/dev/null | head -24
