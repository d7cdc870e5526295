#!/bin/bash
mkdir -p gpurun_out
python -m step_b200.build > gpurun_out/build.log 2>&1 || { cat gpurun_out/build.log; exit 1; }
for n in 1 2 3 4; do
python bench.py --steps 30 --warmup 5 --skip-cpu --inflight $n > gpurun_out/z_bench_$n.log 2>&1; python - <<PY
import json
d=json.loads(open('gpurun_out/z_bench_$n.log').read().strip().splitlines()[-1])
print("inflight $n: clips/s", d["value"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["achieved"])
PY
done
timeout 600 ncu --kernel-name-base demangled -k regex:step:: --metrics gpu__time_duration.sum --clock-control none -s 300 -c 300 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 3 --skip-cpu --no-graph --inflight 1 > gpurun_out/l.log 2>&1
python tools/launch_summary.py gpurun_out/launches_r1.csv | head -20
