#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_bench.log 2>&1
tail -c 600 gpurun_out/ncu_bench.log
