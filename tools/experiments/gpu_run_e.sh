#!/bin/bash
mkdir -p gpurun_out
python -m step_b200.build > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
timeout 300 python -m pytest tests/test_gpu_conv.py -q -m gpu --tb=short -x -k "maxpool" 2>&1 | tail -3
timeout 600 ncu --kernel-name-base demangled -k 'regex:clip_to_s2d|maxpool3d_kernel|mean_mid|linear_splitk' --set full --clock-control none --import-source on -s 12 -c 10 -o gpurun_out/prof_misc2 -f python bench.py --steps 1 --warmup 3 --skip-cpu --no-graph --inflight 1 > gpurun_out/p4.log 2>&1
python bench.py --steps 30 --warmup 5 --skip-cpu > gpurun_out/y_bench.log 2>&1; python - <<'PY'
import json
d=json.loads(open('gpurun_out/y_bench.log').read().strip().splitlines()[-1])
print("clips/s", d["value"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["achieved"], d["roofline"]["ms_per_step_in_kernel"])
PY
