#!/bin/bash
# round 2: how many independent batches to keep in flight (bench.py --inflight)
for n in 1 2 3 4 5 6 8; do
  python bench.py --skip-cpu --steps 60 --inflight $n 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('inflight', d['config']['batches_in_flight'], d['value'], d['e2e']['value'], d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done
