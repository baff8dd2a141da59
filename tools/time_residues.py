"""Device time of the fused residue kernels vs the reference's element-wise torch formulation on the same GPU."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from cpfn_b200.spfn import losses_implementation, metric_implementation

dev = torch.device("cuda:0")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return round(float(np.median(ts)), 1)


g = torch.Generator(device="cpu").manual_seed(3)
B, K, n_pts = 16, 28, 512
unit = lambda t: torch.nn.functional.normalize(t, dim=-1)
pt = {"plane_normal": unit(torch.randn(B, K, 3, generator=g)), "plane_center": 0.3 * torch.randn(B, K, generator=g),
      "sphere_center": 0.3 * torch.randn(B, K, 3, generator=g), "sphere_radius_squared": 0.01 + 0.5 * torch.rand(B, K, generator=g),
      "cylinder_axis": unit(torch.randn(B, K, 3, generator=g)), "cylinder_center": 0.3 * torch.randn(B, K, 3, generator=g),
      "cylinder_radius_squared": 0.01 + 0.3 * torch.rand(B, K, generator=g), "cone_apex": 0.5 * torch.randn(B, K, 3, generator=g),
      "cone_axis": unit(torch.randn(B, K, 3, generator=g)), "cone_half_angle": 0.1 + 1.2 * torch.rand(B, K, generator=g)}
pt = {k: v.to(dev) for k, v in pt.items()}
leaf = {k: v.clone().requires_grad_(True) for k, v in pt.items()}
m = torch.stack([torch.randperm(K, generator=g) for _ in range(B)]).to(dev)
Tg = torch.randint(0, 4, (B, K), generator=g).to(dev)
pts = (0.5 * torch.randn(B, K, n_pts, 3, generator=g)).to(dev)
rows = {"residue_loss_B16_K28_512pts": {
    "fused_us": timed(lambda: losses_implementation.compute_residue_loss(pt, m, pts, Tg)),
    "elementwise_torch_us": timed(lambda: losses_implementation.compute_residue_loss(leaf, m, pts, Tg)),
    "algorithmic_bytes": B * K * n_pts * (12 + 16)}}
for N in (8192, 131072, 1 << 20):
    P = (0.5 * torch.randn(1, N, 3, generator=g)).to(dev)
    pt1 = {k: v[:1].contiguous() for k, v in pt.items()}
    def reference_formulation():
        with torch.no_grad():
            r = metric_implementation.get_residual_loss(pt1, m[:1], P.unsqueeze(1).expand(1, K, N, 3), torch.gather(Tg[:1], 1, m[:1]))
            return ((r.min(dim=1).values < 0.01).float().mean(), (r.min(dim=1).values < 0.02).float().mean())
    rows["p_coverage_K28_N%d" % N] = {
        "fused_us": timed(lambda: metric_implementation.compute_P_coverage(P, Tg[:1], m[:1], pt1, [0.01, 0.02])),
        "reference_formulation_on_gpu_us": timed(reference_formulation), "algorithmic_bytes": 12 * N}
for k, v in rows.items():
    v["fused_GBps"] = round(v["algorithmic_bytes"] / v["fused_us"] / 1e3, 1)
    print(k, v, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/residues_timing.json", "w"), indent=1)
