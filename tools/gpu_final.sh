#!/bin/bash
# Final evidence run: tests, smoke, both bench arms, ncu launch list + full captures of the top kernels.
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
timeout -s KILL 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout -s KILL 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 600 gpurun_out/bench.json
timeout -s KILL 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cut -c1-300 gpurun_out/bench_ref.json
export CPFN_BENCH_NO_CPU=1 CPFN_BENCH_NO_GRAPH=1
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_bench.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:fps_cluster -s 1 -c 1 -f -o gpurun_out/fps_cluster python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_full.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:mlp_chain -s 9 -c 9 -f -o gpurun_out/mlp_chain python bench.py --steps 1 --warmup 1 >> gpurun_out/ncu_full.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"tls_pass|ball_query|three_nn" -s 7 -c 7 -f -o gpurun_out/tls_ball_nn python bench.py --steps 1 --warmup 1 >> gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
