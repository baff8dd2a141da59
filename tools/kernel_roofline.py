"""Per-kernel roofline at bandwidth-relevant sizes (SURVEY.md 8d, item iii): the bench config
(B=16, N=8192) is latency-scale for ball_query / three_nn / TLS, so the same kernels are also timed
where the HBM stream dominates.  CUDA events, L2 flushed between iterations, median of 10.
Writes gpurun_out/kernel_roofline.json (copied to profiles/ by hand)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cpfn_b200 import _lib, cuda_ops, fused, merging_utils, sampling_utils, synth  # noqa: E402
from cpfn_b200.spfn import fit, residues  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timeit(fn, iters=10, warm=3):
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts))


def main():
    dev = torch.device("cuda:0")
    res = {}

    def add(name, us, nbytes, note):
        gbs = nbytes / (us * 1e-6) / 1e9
        res[name] = {"us": round(us, 2), "algorithmic_bytes": int(nbytes), "GB_s": round(gbs, 1),
                     "frac_of_measured_hbm_peak": round(gbs / PEAK, 4), "note": note}
        print(name, res[name], flush=True)

    g = torch.Generator(device="cpu").manual_seed(0)
    for (B, N, K) in [(16, 8192, 28), (1, 8192, 24), (128, 8192, 28), (1, 1 << 20, 28), (8, 131072, 28)]:
        P = (torch.rand(B, N, 3, generator=g) * 2 - 1).to(dev)
        X = torch.nn.functional.normalize(torch.randn(B, N, 3, generator=g), dim=2).to(dev)
        W = torch.softmax(torch.randn(B, N, K, generator=g) * 3, dim=2).to(dev)
        us = timeit(lambda: fit.fit_primitives(P, W, X))
        add("tls_fit B=%d N=%d K=%d" % (B, N, K), us, 2 * B * N * (4 * K + 24),
            "4 types fused, model 2*B*N*(4K+24) B (SURVEY 8d); %d fits" % (4 * B * K))
        del P, X, W
    for B in (16, 256):
        P = torch.from_numpy(synth.uniform_cloud(B, 8192, seed=1)).to(dev)
        i1 = cuda_ops.farthest_point_sampling(P[:16].contiguous(), 512) if B > 16 else cuda_ops.farthest_point_sampling(P, 512)
        c = fused.gather_xyz(P[:16].contiguous(), i1).repeat(B // 16, 1, 1).contiguous() if B > 16 else fused.gather_xyz(P, i1)
        us = timeit(lambda: cuda_ops.ball_query(c, P, 0.2, 64))
        add("ball_query B=%d N=8192 S=512 K=64" % B, us, B * (12 * 8192 + 12 * 512 + 4 * 512 * 64), "compulsory bytes B*(12N+12S+4SK)")
        us = timeit(lambda: fused.three_nn_weights(P, c))
        add("three_nn_weights B=%d n=8192 m=512" % B, us, B * (36 * 8192 + 12 * 512), "compulsory bytes B*(36n+12m)")
        if B == 16:
            us = timeit(lambda: cuda_ops.farthest_point_sampling(P, 512))
            add("fps B=16 N=8192 m=512", us, B * 511 * 8192 * 16, "effective bytes B*(m-1)*N*16 (data is on-chip)")
    P = torch.from_numpy(synth.uniform_cloud(128, 8192, seed=2)).to(dev)
    us = timeit(lambda: cuda_ops.farthest_point_sampling(P, 512))
    add("fps B=128 N=8192 m=512", us, 128 * 511 * 8192 * 16, "effective bytes; one CTA per cloud (B*C > #SMs)")
    del P
    # ---- the widened path ----
    for N in (131072, 1 << 20):
        hr = torch.from_numpy(synth.shape_cloud(N, 77)[0].astype(np.float32)).to(dev)
        for S in (1, 32):
            seeds = hr[:: N // S][:S].contiguous()
            us = timeit(lambda: sampling_utils.extract_patches(hr, seeds, 8192, return_distances=True))
            add("extract_patches N=%d seeds=%d k=8192" % (N, S), us, 12 * N + S * 8192 * 8, "12*N + S*k*8 B (SURVEY 8d)")
    Ng, nb, Np, Kl, Kg = 131072, 32, 8192, 21, 28
    hr = torch.from_numpy(synth.shape_cloud(Ng, 99)[0].astype(np.float32)).to(dev)
    idx = sampling_utils.extract_patches(hr, hr[:: Ng // nb][:nb].contiguous(), Np)
    Wp = torch.softmax(4 * torch.randn(nb, Np, Kl, generator=g).to(dev), dim=2)
    Sl = torch.nn.functional.one_hot(torch.randint(Kg, (Ng,), generator=g), Kg).float().to(dev)
    inverse = merging_utils.inverse_index(idx, Ng)
    us = timeit(lambda: merging_utils.similarity_soft(Sl, Wp, idx, inverse=inverse))
    add("merge_similarity N=131072 32x8192 Kl=21 Kg=28", us, nb * Np * Kl * 4 + nb * Np * 4 + Ng * Kg * 4,
        "W + patch indices + object labels read once; dense reference formulation = 128 GFLOP")
    B, K, n_pts = 16, 28, 512
    unit = lambda t_: torch.nn.functional.normalize(t_, dim=-1)
    pt = {"plane_normal": unit(torch.randn(B, K, 3, generator=g)), "plane_center": 0.3 * torch.randn(B, K, generator=g),
          "sphere_center": 0.3 * torch.randn(B, K, 3, generator=g), "sphere_radius_squared": 0.01 + 0.5 * torch.rand(B, K, generator=g),
          "cylinder_axis": unit(torch.randn(B, K, 3, generator=g)), "cylinder_center": 0.3 * torch.randn(B, K, 3, generator=g),
          "cylinder_radius_squared": 0.01 + 0.3 * torch.rand(B, K, generator=g), "cone_apex": 0.5 * torch.randn(B, K, 3, generator=g),
          "cone_axis": unit(torch.randn(B, K, 3, generator=g)), "cone_half_angle": 0.1 + 1.2 * torch.rand(B, K, generator=g)}
    pt = {k: v.to(dev) for k, v in pt.items()}
    match = torch.stack([torch.randperm(K, generator=g) for _ in range(B)]).to(dev)
    pts = (0.5 * torch.randn(B, K, n_pts, 3, generator=g)).to(dev)
    us = timeit(lambda: residues.residues(pt, match, pts))
    add("primitive_residues B=16 K=28 512 pts 4 types", us, B * K * n_pts * (12 + 16), "12 B in + 16 B out per point")
    cloud = torch.from_numpy(synth.shape_cloud(131072, 5)[0].astype(np.float32)).to(dev)
    out = torch.empty(8192, dtype=torch.int32, device=dev)
    lib = _lib.lib()
    ws = torch.empty(lib.cpfn_fps_dense_workspace_bytes(), dtype=torch.uint8, device=dev)
    us = timeit(lambda: _lib.check(lib.cpfn_fps_dense(cloud.data_ptr(), 131072, None, None, 0, 0, 8192, out.data_ptr(), ws.data_ptr(),
                                                      ws.numel(), torch.cuda.current_stream(dev).cuda_stream), "fps_dense"), iters=3, warm=1)
    add("fps_dense N=131072 m=8192", us, 8191 * 131072 * 16, "effective bytes (m-1)*N*16; data is register resident")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump({"hbm_peak_gbs": PEAK, "kernels": res}, open(os.path.join(ROOT, "gpurun_out", "kernel_roofline.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
