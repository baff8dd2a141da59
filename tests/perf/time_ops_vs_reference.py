"""Per-op device timings at GlobalSPFN sizes (B=16, N=8192): this library vs the
reference kernels (oracle/_ref) on the same GPU.  CUDA events, L2 flushed between
iterations.  Diagnostic tool, not the bench contract (see bench.py); lives under tests/ because it
loads the reference extension through oracle/ (test infrastructure)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cpfn_b200 import cuda_ops, synth  # noqa: E402
from oracle import build_ref  # noqa: E402


def timeit(fn, iters=10, warm=3):
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts)), float(np.min(ts))


def main():
    B = int(os.environ.get("B", 16))
    ref = build_ref.load_module()
    dev = torch.device("cuda:0")
    P = torch.from_numpy(synth.shape_batch(B, 8192, seed=1235)[0]).to(dev)
    res = {}

    def both(name, ours, theirs):
        res[name] = {"ours_us": timeit(ours)}
        if ref is not None:
            res[name]["ref_us"] = timeit(theirs)
        print(name, res[name], flush=True)

    both("fps_8192_512", lambda: cuda_ops.farthest_point_sampling(P, 512),
         lambda: ref.farthest_point_sampling(P, 512))
    i1 = cuda_ops.farthest_point_sampling(P, 512)
    c1 = torch.gather(P, 1, i1.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    both("fps_512_128", lambda: cuda_ops.farthest_point_sampling(c1, 128),
         lambda: ref.farthest_point_sampling(c1, 128))
    i2 = cuda_ops.farthest_point_sampling(c1, 128)
    c2 = torch.gather(c1, 1, i2.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    both("ball_sa1", lambda: cuda_ops.ball_query(c1, P, 0.2, 64), lambda: ref.ball_query(c1, P, 0.2, 64))
    both("ball_sa2", lambda: cuda_ops.ball_query(c2, c1, 0.4, 64), lambda: ref.ball_query(c2, c1, 0.4, 64))
    both("three_nn_fp3", lambda: cuda_ops.three_nn(P, c1), lambda: ref.three_nn(P, c1))
    both("three_nn_fp2", lambda: cuda_ops.three_nn(c1, c2), lambda: ref.three_nn(c1, c2))
    d2, idx = cuda_ops.three_nn(P, c1)
    w = torch.rand(B, 8192, 3, device=dev)
    f = torch.randn(B, 128, 512, device=dev)
    both("tws_fp3", lambda: cuda_ops.three_weighted_sum(f, idx, w), lambda: ref.three_weighted_sum(f, idx, w))
    g = torch.randn(B, 128, 8192, device=dev)
    both("tws_grad_fp3", lambda: cuda_ops.three_weighted_sum_grad(g, idx, w, 512),
         lambda: ref.three_weighted_sum_grad(g, idx, w, 512))
    gi = cuda_ops.ball_query(c2, c1, 0.4, 64)
    f1 = torch.randn(B, 128, 512, device=dev)
    both("group_sa2", lambda: cuda_ops.group_points(f1, gi), lambda: ref.group_points(f1, gi))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "time_ops.json"), "w") as fjs:
        json.dump(res, fjs, indent=1)


if __name__ == "__main__":
    main()
