#!/bin/bash
# Round-2 multi-GPU evidence on N GPUs of one box: the patch-sharded cascade / training tests that need N ranks, then
# the bench line at N (weak scaling of the headline metric + the cascade and training sub-records).
#   bash tools/gpu_r2_multi.sh N [skip_tests]
set -u
N=${1:-2}
mkdir -p gpurun_out
if [ -z "${2:-}" ]; then
  timeout 900 python -m pytest tests/test_gpu_cascade_nccl.py tests/test_gpu_training.py -q -x > gpurun_out/r2_pytest_n${N}.log 2>&1
  echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_n${N}.log
fi
PORT=$((29500 + N))
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
  bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/r2_bench_n${N}.json 2> gpurun_out/r2_bench_n${N}.err
echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_bench_n${N}.json").read().strip().splitlines()[-1])
print("N=%d value %.1f M pts/s  %.4f ms/step  e2e %.1f M pts/s" % (d["n_gpus"], d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6))
print("cascade", json.dumps(d.get("cascade"))[:900])
print("training", json.dumps(d.get("training"))[:500])
PY
