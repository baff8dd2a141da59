#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_gpu_merging.py -x -q 2>&1 | tail -30 > gpurun_out/pytest_merging.log; tail -30 gpurun_out/pytest_merging.log
timeout -s KILL 300 python tools/time_merging.py 2>&1 | tee gpurun_out/time_merging.log | tail -20
