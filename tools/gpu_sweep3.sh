#!/bin/bash
export CPFN_BENCH_NO_CPU=1
for v in 8 4 2; do
  echo "== FPS_CLUSTER=$v"
  CPFN_FPS_CLUSTER=$v timeout -s KILL 300 python bench.py --steps 10 --warmup 3 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln); print(d['ms_per_step'], d['breakdown_us']['farthest_point_sampling'], d['roofline']['kernel_us'])
"
done
