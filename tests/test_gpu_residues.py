"""GPU parity of the fused residue kernels through the C ABI: compute_residue_loss,
get_residual_loss and compute_P_coverage vs the goldens of the unmodified reference and the numpy oracle
(fp32, tolerance 2e-5 relative: acos / sin / sqrt come from different math libraries), the autograd path vs the
kernel path, and both at training / evaluation sizes against the element-wise torch formulation on the GPU."""
import os

import numpy as np
import pytest
import torch

from cpfn_b200.spfn import losses_implementation, metric_implementation, residues
from oracle import residues as ores
from tests.golden import cases

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_residues.npz"))
TOL = 2e-5


def _close(a, b, tol=TOL):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() <= tol * max(1.0, np.abs(b).max())


def _case(dev):
    params, matching, points, T_gt, P = cases.residue_case()
    return ({k: torch.from_numpy(v).to(dev) for k, v in params.items()},) + tuple(
        torch.from_numpy(a).to(dev) for a in (matching, points, T_gt, P))


def test_residue_loss_matches_reference(cuda_dev):
    pt, m, pts, Tg, P = _case(cuda_dev)
    loss, per_point = losses_implementation.compute_residue_loss(pt, m, pts, Tg)
    assert per_point.shape == (2, 6, 96, 4) and loss.shape == (2, 6)
    assert _close(per_point.cpu().numpy(), GOLD["residue_per_point"]) and _close(loss.cpu().numpy(), GOLD["residue_loss"])
    params, matching, points, T_gt, _ = cases.residue_case()
    o_loss, o_pp = ores.compute_residue_loss(params, matching, points, T_gt)
    assert _close(per_point.cpu().numpy(), o_pp) and _close(loss.cpu().numpy(), o_loss)
    loss2, pp2 = losses_implementation.compute_residue_loss(pt, m, pts, Tg.clamp(max=1), classes=['cone', 'plane'])
    assert _close(pp2.cpu().numpy(), GOLD["residue_per_point_cone_plane"])
    assert _close(loss2.cpu().numpy(), GOLD["residue_loss_cone_plane"])
    assert _close(metric_implementation.get_residual_loss(pt, m, pts, Tg).cpu().numpy(), GOLD["residual"])
    with pytest.raises(NotImplementedError):
        losses_implementation.compute_residue_loss(pt, m, pts, Tg, classes=['torus'])
    with pytest.raises(RuntimeError):
        residues.residues({k: v.cpu() for k, v in pt.items()}, m.cpu(), pts.cpu())


def test_p_coverage_matches_reference(cuda_dev):
    pt, m, pts, Tg, P = _case(cuda_dev)
    for eps in (0.05, 0.2):
        got = metric_implementation.compute_P_coverage(P, Tg, m, pt, eps)
        assert got.shape == (2,) and np.abs(got.cpu().numpy() - GOLD["p_coverage_%g" % eps]).max() <= 1.5 / P.shape[1]
    both = metric_implementation.compute_P_coverage(P, Tg, m, pt, [0.05, 0.2])
    assert both.shape == (2, 2) and np.abs(both[:, 1].cpu().numpy() - GOLD["p_coverage_0.2"]).max() <= 1.5 / P.shape[1]


def test_autograd_path_agrees_and_differentiates(cuda_dev):
    pt, m, pts, Tg, P = _case(cuda_dev)
    ref_loss, ref_pp = losses_implementation.compute_residue_loss(pt, m, pts, Tg)          # kernel
    leaf = {k: v.clone().requires_grad_(True) for k, v in pt.items()}
    loss, pp = losses_implementation.compute_residue_loss(leaf, m, pts, Tg)                # element-wise torch
    assert loss.requires_grad and _close(pp.detach().cpu().numpy(), ref_pp.cpu().numpy())
    loss.sum().backward()
    assert all(v.grad is not None and torch.isfinite(v.grad).all() for v in leaf.values())
    with torch.no_grad():                                                                  # no_grad: the kernel again
        again, _ = losses_implementation.compute_residue_loss(leaf, m, pts, Tg)
    assert torch.equal(again, ref_loss)


def test_training_and_evaluation_sizes(cuda_dev):
    g = torch.Generator(device="cpu").manual_seed(3)
    B, K, n_pts, N = 16, 28, 512, 131072
    unit = lambda t: torch.nn.functional.normalize(t, dim=-1)
    pt = {"plane_normal": unit(torch.randn(B, K, 3, generator=g)), "plane_center": 0.3 * torch.randn(B, K, generator=g),
          "sphere_center": 0.3 * torch.randn(B, K, 3, generator=g), "sphere_radius_squared": 0.01 + 0.5 * torch.rand(B, K, generator=g),
          "cylinder_axis": unit(torch.randn(B, K, 3, generator=g)), "cylinder_center": 0.3 * torch.randn(B, K, 3, generator=g),
          "cylinder_radius_squared": 0.01 + 0.3 * torch.rand(B, K, generator=g), "cone_apex": 0.5 * torch.randn(B, K, 3, generator=g),
          "cone_axis": unit(torch.randn(B, K, 3, generator=g)), "cone_half_angle": 0.1 + 1.2 * torch.rand(B, K, generator=g)}
    pt = {k: v.to(cuda_dev) for k, v in pt.items()}
    m = torch.stack([torch.randperm(K, generator=g) for _ in range(B)]).to(cuda_dev)
    Tg = torch.randint(0, 4, (B, K), generator=g).to(cuda_dev)
    pts = (0.5 * torch.randn(B, K, n_pts, 3, generator=g)).to(cuda_dev)
    loss, pp = losses_implementation.compute_residue_loss(pt, m, pts, Tg)
    leaf = {k: v.clone().requires_grad_(True) for k, v in pt.items()}
    t_loss, t_pp = losses_implementation.compute_residue_loss(leaf, m, pts, Tg)            # same formulas, torch kernels
    assert _close(pp.cpu().numpy(), t_pp.detach().cpu().numpy()) and _close(loss.cpu().numpy(), t_loss.detach().cpu().numpy())
    # evaluation: one shape, full-resolution cloud, every primitive against every point
    P = (0.5 * torch.randn(1, N, 3, generator=g)).to(cuda_dev)
    pt1 = {k: v[:1].contiguous() for k, v in pt.items()}
    cov = metric_implementation.compute_P_coverage(P, Tg[:1], m[:1], pt1, [0.01, 0.02, 0.1])
    with torch.no_grad():                                                                  # the reference's formulation
        r = metric_implementation.get_residual_loss(pt1, m[:1], P.unsqueeze(1).expand(1, K, N, 3), torch.gather(Tg[:1], 1, m[:1]))
        rmin = r.min(dim=1).values
    for i, eps in enumerate((0.01, 0.02, 0.1)):
        want = float((rmin < eps).float().mean())
        assert abs(float(cov[0, i]) - want) <= 1e-4, (eps, float(cov[0, i]), want)
    assert 0.0 < float(cov[0, 0]) < float(cov[0, 2]) < 1.0
