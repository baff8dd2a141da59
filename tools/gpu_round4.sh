#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_index_ops.py -x -q 2>&1 | tail -6
timeout -s KILL 600 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/pytest_gpu.log; tail -2 gpurun_out/pytest_gpu.log
CPFN_BENCH_NO_CPU=1 timeout -s KILL 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; python -c "
import json
d = json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['breakdown_us'], d['roofline']['kernel_us'])
"; tail -3 gpurun_out/bench.err
