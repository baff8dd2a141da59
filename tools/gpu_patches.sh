#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_patches.py -x -q 2>&1 | tail -30 > gpurun_out/pytest_patches.log; tail -30 gpurun_out/pytest_patches.log
timeout -s KILL 200 python tools/time_patches.py 2>&1 | tee gpurun_out/time_patches.log | tail -20
