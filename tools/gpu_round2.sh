#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_mlp_chain.py -x -q 2>&1 | tail -40 > gpurun_out/pytest_chain.log; tail -30 gpurun_out/pytest_chain.log
if grep -q "passed" gpurun_out/pytest_chain.log && ! grep -q "failed" gpurun_out/pytest_chain.log; then
  timeout -s KILL 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log; tail -15 gpurun_out/pytest_gpu.log
  timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
  timeout -s KILL 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 2500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
fi
