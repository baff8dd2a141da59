#!/bin/bash
# Round-2 evidence run on one B200: full GPU test suite, smoke, the bench (both arms), the ncu launch list and one
# `--set full` capture of an eager step, the chains' phase profile, the per-kernel roofline table and the f3 timings.
# Everything lands in gpurun_out/ (copied into profiles/ by hand).
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2_pytest_gpu.log
tail -3 gpurun_out/r2_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err; echo "reference arm rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none --launch-skip 143 --launch-count 23 --csv --log-file gpurun_out/r2_launches.csv python tools/profile_step.py > /dev/null 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --launch-skip 143 --launch-count 23 -f -o gpurun_out/r2_step_final python tools/profile_step.py > gpurun_out/r2_ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 300 python tools/chain_profile.py > gpurun_out/r2_chain_profile.txt 2>&1; tail -10 gpurun_out/r2_chain_profile.txt
timeout 600 python tools/kernel_roofline.py > gpurun_out/r2_kernel_roofline.log 2>&1; echo "roofline rc=$?"
timeout 300 python tools/time_seg.py > gpurun_out/r2_seg.log 2>&1; echo "seg rc=$?"; cat gpurun_out/r2_seg.md 2>/dev/null
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench_n1.json", "gpurun_out/r2_bench_reference_arm.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d.get("value"), d.get("ms_per_step"), d.get("e2e", {}).get("value"), d.get("roofline", {}).get("frac"), d.get("cpu_baseline", {}).get("value"))
    except Exception as e:
        print(f, "unreadable", e)
PY
