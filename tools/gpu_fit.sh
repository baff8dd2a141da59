#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_fitters.py -x -q 2>&1 | tail -25
