"""Times the segmentation glue -- hungarian_matching + compute_miou_loss of cpfn_b200.spfn.seg --
against the unmodified reference functions (staged under baseline/_ref) on the same GPU tensors.  CUDA events around
the calls with a synchronise on both sides (the reference synchronises by itself: it copies the cost matrix to the
host for scipy).  Writes gpurun_out/r2_seg.md."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cpfn_b200 import synth                                   # noqa: E402
from cpfn_b200.spfn import seg                                # noqa: E402
from oracle import ref_runtime                                # noqa: E402


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e6)
    return float(np.median(ts))


def main():
    dev = torch.device("cuda:0")
    ref = ref_runtime.load_spfn() if ref_runtime.available() else None
    rows = []
    for (B, N, K) in [(16, 8192, 28), (16, 8192, 21), (1, 131072, 28), (1, 1 << 20, 28)]:
        g = torch.Generator().manual_seed(B + N)
        I = torch.randint(-1, K - 3, (B, N), generator=g).to(dev)
        W = torch.softmax(3 * torch.randn(B, N, K, generator=g), dim=2).to(dev)

        def ours():
            m = seg.hungarian_matching(W, I)
            return seg.compute_miou_loss(W, I, m)

        t_ours = timeit(ours)
        t_ref = None
        if ref is not None:
            L = ref

            def theirs():
                m = L.hungarian_matching(W, I)
                return L.compute_miou_loss(W, I, m)

            t_ref = timeit(theirs, iters=5, warm=1)
            a, b = ours(), theirs()
            assert torch.allclose(a[0], b[0], atol=1e-5), (a[0], b[0])
        nbytes = 2 * 4 * B * N * (K + 1)
        rows.append((B, N, K, t_ours, t_ref, nbytes / t_ours / 1e3))
        print(rows[-1], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "r2_seg.md"), "w") as f:
        f.write("# hungarian_matching + compute_miou_loss (row f3), wall clock per call incl. synchronise, median\n\n")
        f.write("| B | N | K | this package (us) | reference functions, same GPU (us) | GB/s of 2*4*B*N*(K+1) B |\n|---|---|---|---|---|---|\n")
        for r in rows:
            f.write("| %d | %d | %d | %.1f | %s | %.1f |\n" % (r[0], r[1], r[2], r[3], "%.1f" % r[4] if r[4] else "n/a", r[5]))


if __name__ == "__main__":
    main()
