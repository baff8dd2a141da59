"""SPFN/geometry_utils.py API: ``weighted_plane_fitting`` (:74-84), ``weighted_sphere_fitting``
(:209-223), ``guarded_matrix_solve_ls`` (:121-142), ``compute_consistent_plane_frame`` (:8-27).
The per-point sums run on the CUDA moment kernels where the shapes allow (3-D points); the 2-D
circle fit and the generic least squares on explicit [B',N,D] rows are thin torch compositions
(inside ``compute_parameters`` the cylinder's circle fit never builds those rows, see csrc/tls.cu)."""
import torch

from . import _reference, _train

compute_consistent_plane_frame = _train.compute_consistent_plane_frame


def weighted_plane_fitting(P, W, division_eps=1e-10):
    """P [B',N,3], W [B',N] -> (n [B',3], c [B']).  Differentiable w.r.t. W and P."""
    M = _train.weighted_moments(W.unsqueeze(2), P.detach(), P).squeeze(1)
    Sw = M[..., 0]
    denom = torch.clamp(Sw, min=division_eps)
    R1 = M[..., _train._X1:_train._X1 + 3]
    R2 = _train._sym3(M[..., _train._X2:_train._X2 + 6])
    mu = R1 / denom[..., None]
    S = R2 - mu[..., :, None] * R1[..., None, :] - R1[..., :, None] * mu[..., None, :] \
        + Sw[..., None, None] * mu[..., :, None] * mu[..., None, :]
    n = _train.svd_v_last_column(S)
    return n.to(torch.float32), (n * mu).sum(-1).to(torch.float32)


def guarded_matrix_solve_ls(A, b, W, condition_number_cap=1e5, sqrt_eps=1e-10, ls_l2_regularizer=1e-8):
    """A [B',N,D], b [B',N,1], W [B',N] -> x [B',D]."""
    w = torch.clamp(W, min=sqrt_eps).unsqueeze(2).double()
    Ad, bd = A.double(), b.double()
    AtA = (Ad * w).transpose(1, 2) @ Ad
    Atb = ((Ad * w).transpose(1, 2) @ bd).squeeze(2)
    return _train.guarded_solve(AtA, Atb, condition_number_cap, ls_l2_regularizer).to(A.dtype)


def weighted_sphere_fitting(P, W, division_eps=1e-10):
    """P [B',N,D] (D = 2 or 3), W [B',N] -> (center [B',D], radius_squared [B'])."""
    D = P.shape[2]
    if D == 3 and not P.requires_grad and P.is_cuda:
        X0 = torch.zeros_like(P)
        Wc = torch.clamp(W, min=division_eps)
        M = _train.weighted_moments(W.unsqueeze(2), P, X0).squeeze(1)
        Mc = _train.weighted_moments(Wc.unsqueeze(2), P, X0).squeeze(1)
        c, r2 = _train._sphere_from_moments(M[..., 1:4], _train._sym3(M[..., 4:10]), None, M[..., 0], Mc[..., 1:4],
                                            _train._sym3(Mc[..., 4:10]), _train._t3(Mc[..., 10:20]), Mc[..., 0])
        return c.to(torch.float32), r2.to(torch.float32)
    W_sum = torch.sum(W, dim=1)
    denom = torch.clamp(W_sum, min=division_eps)
    P_sqr = torch.sum(P ** 2, dim=2)
    b = ((torch.sum(W * P_sqr, dim=1) / denom).unsqueeze(1) - P_sqr).unsqueeze(2)
    A = 2 * ((torch.sum(W.unsqueeze(2) * P, dim=1) / denom.unsqueeze(1)).unsqueeze(1) - P)
    center = guarded_matrix_solve_ls(A, b, W)
    r2 = torch.sum(W * torch.sum((P - center.unsqueeze(1)) ** 2, dim=2), dim=1) / denom
    return center, r2


__getattr__ = _reference.forwarder(globals(), "geometry_utils", ('compute_consistent_plane_frame', 'weighted_plane_fitting', 'guarded_matrix_solve_ls', 'weighted_sphere_fitting'))
