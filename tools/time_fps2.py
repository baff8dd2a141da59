import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpfn_b200 import cuda_ops, synth
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for C in (4, 8, 2):
    os.environ["CPFN_FPS_CLUSTER"] = str(C)
    for N in (2048, 4096, 8192, 16384):
        P = torch.from_numpy(synth.shape_batch(16, N, seed=1234)[0]).to(dev)
        cuda_ops.farthest_point_sampling(P, 512); torch.cuda.synchronize()
        ts = []
        for _ in range(6):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); cuda_ops.farthest_point_sampling(P, 512); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        print("C", C, "N", N, "ppt", N // C // 256, "us %.1f" % np.median(ts), "ns/round %.0f" % (np.median(ts) * 1e3 / 511), flush=True)
