#!/bin/bash
mkdir -p gpurun_out
python -m step_b200.build > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
timeout 300 python -m pytest tests/test_gpu_conv.py tests/test_gpu_pipeline.py -q -m gpu --tb=short -x -k "s2d or stem or pipeline_fp16 or c1" 2>&1 | tail -3
echo "== default"; python tools/conv_bench.py loc_res loc_nores 2>&1 | tail -2
echo "== MH=2"; STEP_B200_MH=2 python tools/conv_bench.py loc_res loc_nores 2>&1 | tail -2
echo "== STAGES=2"; STEP_B200_STAGES=2 python tools/conv_bench.py loc_res loc_nores 2>&1 | tail -2
echo "== CONV=1"; STEP_B200_CONV=1 python tools/conv_bench.py loc_res loc_nores 2>&1 | tail -2
timeout 300 ncu --kernel-name-base demangled -k 'regex:conv_umma_persist_kernel' --set full --clock-control none --import-source on -s 3 -c 1 -o gpurun_out/prof_locres -f python tools/conv_bench.py loc_res > gpurun_out/f1.log 2>&1
timeout 300 ncu --kernel-name-base demangled -k 'regex:conv_umma_persist_kernel' --set full --clock-control none --import-source on -s 3 -c 1 -o gpurun_out/prof_locnores -f python tools/conv_bench.py loc_nores > gpurun_out/f2.log 2>&1
python bench.py --steps 30 --warmup 5 --skip-cpu > gpurun_out/y_bench.log 2>&1; python - <<'PY'
import json
d=json.loads(open('gpurun_out/y_bench.log').read().strip().splitlines()[-1])
print("clips/s", d["value"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["achieved"], d["roofline"]["ms_per_step_in_kernel"])
PY
