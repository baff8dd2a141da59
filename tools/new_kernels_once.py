"""One call of the widened-path kernels (merge, residues, patch selection) for an ncu --set full capture."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cpfn_b200 import merging_utils, sampling_utils, synth
from cpfn_b200.spfn import losses_implementation, metric_implementation
dev = torch.device("cuda:0")
Ng, nb, Np, Kl, Kg = 131072, 32, 8192, 21, 28
g = torch.Generator(device="cpu").manual_seed(5)
P = torch.from_numpy(synth.shape_cloud(Ng, 99)[0]).to(dev)
seeds = P[:: Ng // nb][:nb].contiguous()
for rep in range(2):
    idx = sampling_utils.extract_patches(P, seeds, Np)
    W = torch.softmax(4 * torch.randn(nb, Np, Kl, generator=g).to(dev), dim=2)
    S = torch.nn.functional.one_hot(torch.randint(Kg, (Ng,), generator=g), Kg).float().to(dev)
    X = torch.nn.functional.normalize(torch.randn(nb, Np, 3, generator=g), dim=2).to(dev)
    T = torch.randn(nb, Np, 4, generator=g).to(dev)
    on = torch.nn.functional.normalize(torch.randn(Ng, 3, generator=g), dim=1).to(dev)
    ot = torch.randn(Ng, 4, generator=g).to(dev)
    merging_utils.merge_shape(W, X, T, idx, S, on, ot)
    B, K = 16, 28
    unit = lambda t: torch.nn.functional.normalize(t, dim=-1)
    pt = {"plane_normal": unit(torch.randn(B, K, 3, generator=g)), "plane_center": 0.3 * torch.randn(B, K, generator=g),
          "sphere_center": 0.3 * torch.randn(B, K, 3, generator=g), "sphere_radius_squared": 0.01 + 0.5 * torch.rand(B, K, generator=g),
          "cylinder_axis": unit(torch.randn(B, K, 3, generator=g)), "cylinder_center": 0.3 * torch.randn(B, K, 3, generator=g),
          "cylinder_radius_squared": 0.01 + 0.3 * torch.rand(B, K, generator=g), "cone_apex": 0.5 * torch.randn(B, K, 3, generator=g),
          "cone_axis": unit(torch.randn(B, K, 3, generator=g)), "cone_half_angle": 0.1 + 1.2 * torch.rand(B, K, generator=g)}
    pt = {k: v.to(dev) for k, v in pt.items()}
    m = torch.stack([torch.randperm(K, generator=g) for _ in range(B)]).to(dev)
    Tg = torch.randint(0, 4, (B, K), generator=g).to(dev)
    pts = (0.5 * torch.randn(B, K, 512, 3, generator=g)).to(dev)
    losses_implementation.compute_residue_loss(pt, m, pts, Tg)
    metric_implementation.compute_P_coverage(P[None], Tg[:1], m[:1], {k: v[:1].contiguous() for k, v in pt.items()}, [0.01, 0.02])
torch.cuda.synchronize()
