#!/bin/bash
mkdir -p gpurun_out
python -m step_b200.build > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_pipeline.py -q -m gpu --tb=short -x 2>&1 | tail -3
echo "== deep"; CB_AMODE=3 python tools/conv_bench.py 4c_b1b 4e_b1b 2>&1 | tail -2
echo "== DEEP=0"; STEP_B200_DEEP=0 CB_AMODE=3 python tools/conv_bench.py 4c_b1b 4e_b1b 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 --skip-cpu 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['roofline']['frac'])"
