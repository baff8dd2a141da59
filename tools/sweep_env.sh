#!/bin/bash
# A/B of environment knobs on the bench's device-resident throughput (no CPU / reference / cascade / training legs).
run() {
  env "$@" CPFN_BENCH_NO_CPU=1 CPFN_BENCH_NO_GREF=1 CPFN_BENCH_NO_CASCADE=1 CPFN_BENCH_NO_TRAIN=1 python bench.py --steps 60 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('%-40s pipelined %.4f  sequential %.4f  e2e %.4f ms  fps %.1f us' % ('$*', d['ms_per_step'], d['sequential_ms_per_step'], 16*8192/d['e2e']['value']*1e3, d['roofline']['kernel_us']))"
}
run A=0
run CPFN_TILE_LAYERWISE=32
run A=1
run CPFN_TILE_LAYERWISE=32
run CPFN_LANES=8
run CPFN_LANES=10
run A=2
