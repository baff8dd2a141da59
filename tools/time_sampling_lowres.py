"""Device time of cpfn_fps_dense at the preprocessing sizes; the per-round work of the reference's numba loop
(distance sweep, running minimum, arg-max -- plain numpy here) is timed beside it on a bounded number of rounds."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from cpfn_b200 import _lib, synth

dev = torch.device("cuda:0")
rows = []
for N in (131072, 1 << 20):
    P = synth.shape_cloud(N, 5)[0].astype(np.float32)
    Pd = torch.from_numpy(P).to(dev)
    out = torch.empty(8192, dtype=torch.int32, device=dev)
    lib = _lib.lib()
    ws = torch.empty(lib.cpfn_fps_dense_workspace_bytes(), dtype=torch.uint8, device=dev)
    def run():
        _lib.check(lib.cpfn_fps_dense(Pd.data_ptr(), N, None, None, 0, 0, 8192, out.data_ptr(), ws.data_ptr(), ws.numel(),
                                      torch.cuda.current_stream(dev).cuda_stream), "fps_dense")
    run(); torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = float(np.median(ts))
    rounds = 64
    running, pick = np.full(N, 1e6), 0
    t0 = time.perf_counter()
    for _ in range(rounds):
        running = np.minimum(running, np.sqrt(np.sum((P - P[pick]) ** 2, axis=1)))
        pick = int(np.argmax(running))
    cpu = time.perf_counter() - t0
    rows.append({"N": N, "samples": 8192, "ms": round(ms, 2), "us_per_round": round(ms * 1e3 / 8192, 2),
                 "effective_GBps": round(8191 * N * 16 / ms / 1e6, 1),
                 "numpy_ms_per_round": round(cpu / rounds * 1e3, 2), "numpy_s_extrapolated_8192": round(cpu / rounds * 8192, 1)})
    print(rows[-1], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/sampling_lowres_timing.json", "w"), indent=1)
