"""Device time of cpfn_extract_patches (CUDA events, L2 flushed between repetitions) + the numpy reference
block timed beside it on the host, for the sizes of BASELINE configs 3 / 5."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from cpfn_b200 import sampling_utils, synth

dev = torch.device("cuda:0")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
peaks = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {}
rows = []
for N, S in ((131072, 1), (131072, 32), (1 << 20, 1), (1 << 20, 32)):
    hr = synth.shape_cloud(N, 77)[0].astype(np.float32)
    hr_d = torch.from_numpy(hr).to(dev)
    seeds_d = hr_d[:: N // S][:S].contiguous()
    for _ in range(3):
        sampling_utils.extract_patches(hr_d, seeds_d, 8192, return_distances=True)
    ts = []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        sampling_utils.extract_patches(hr_d, seeds_d, 8192, return_distances=True)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    us = float(np.median(ts))
    alg = 12 * N + S * 8192 * 8
    t0 = time.perf_counter()
    seed = hr[0]
    d = np.linalg.norm(seed[None] - hr, axis=1)
    np.argsort(d)[:8192]
    np.sort(d)[:8192]
    cpu_us = (time.perf_counter() - t0) * 1e6
    rows.append({"N": N, "seeds": S, "us": round(us, 1), "us_per_seed": round(us / S, 2),
                 "algorithmic_bytes": alg, "GBps": round(alg / us / 1e3, 1), "numpy_us_per_seed": round(cpu_us, 0)})
    print(rows[-1], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/patches_timing.json", "w"), indent=1)
