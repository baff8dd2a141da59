#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_mlp_chain.py tests/test_gpu_network.py -x -q 2>&1 | tail -5
export CPFN_BENCH_NO_CPU=1
for cfg in "128 128" "64 128" "128 64" "64 64"; do
  set -- $cfg
  echo "== SA2=$1 HEAD=$2"
  CPFN_TILE_SA2=$1 CPFN_TILE_HEAD=$2 timeout -s KILL 300 python bench.py --steps 10 --warmup 3 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln); print(d['ms_per_step'], d['breakdown_us'])
    else: print(ln[:200])
"
done
