#!/bin/bash
# GPU bring-up suite: runs the -m gpu tests in risk order, each under its own timeout so a hung
# kernel cannot eat the whole lease.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name" ; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n 12 gpurun_out/$name.log; }
run t1_ops      python -m pytest tests/test_gpu_ops.py -q -m gpu -x --tb=short
run t2_simt     python -m pytest tests/test_gpu_conv.py -q -m gpu -k "simt_fp32" --tb=short
run t3_pipe32   python -m pytest tests/test_gpu_pipeline.py -q -m gpu -k "fp32" --tb=short
run t4_tma      python -m pytest tests/test_gpu_conv.py -q -m gpu -k "tma_tile" --tb=short
run t5_umma     python -m pytest tests/test_gpu_conv.py -q -m gpu -k "umma or stem" --tb=short
run t6_pipe16   python -m pytest tests/test_gpu_pipeline.py -q -m gpu -k "fp16 or driver" --tb=short
run t7_smoke    python __graft_entry__.py smoke
run t8_bench    python bench.py --steps 5 --warmup 3
