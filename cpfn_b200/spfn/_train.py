"""Differentiable (training) path of the SPFN fitters.

The O(B*N*K) work -- every sum over the points the reference's fitters take -- is ONE linear map,
the raw weighted moments M[b,k,:] = sum_n w[b,n,k] * psi(P[b,n], X[b,n]) (psi: monomials up to third
order, 32 features), computed by the CUDA kernel ``cpfn_weighted_moments`` with fp64 accumulation;
its backward (dW, dX) is ``cpfn_weighted_moments_grad``.  Everything after that is algebra on
[B,K,3,3]-sized tensors in float64: the eigen-decompositions and the linear solves are the batched one-thread-per-
matrix kernels of csrc/small_linalg.cu (``cpfn_sym_eigh_small``, ``cpfn_small_solve``; no cuSOLVER / MAGMA call and no
host synchronisation on the path), the identities around them element-wise torch so that autograd differentiates them:
  * ``svd_v_last_column``: forward = eigenvector of the eigenvalue smallest in magnitude (what
    ``torch.svd(M)[2][:, :, -1]`` is for a symmetric M); backward = the reference's analytic formula
    with its guarded 1/(s_i^2 - s_j^2) matrix (SPFN/differentiable_tls.py:8-17, 45-53, 123-143);
  * guarded least squares on the normal equations (SPFN/geometry_utils.py:121-142): condition-number
    mask from detached singular values, ridge 1e-8, ``torch.linalg.solve``;
  * plane / sphere / cylinder / cone parameters expressed through the moments (the same identities
    the inference kernels use, csrc/tls.cu), cone half-angle through an element-wise torch
    composition on [B,N,K] (SPFN/cone_fitter.py:24-35).
P receives no gradient (as in the reference); W and X do.  There is no CPU path: CPU tensors raise.
"""
import math

import torch

from .. import _lib, cuda_ops

NF = 32
# feature offsets inside psi
_P1, _P2, _P3, _X1, _X2, _XP = 1, 4, 10, 20, 23, 29
_SYM6 = ((0, 1, 2), (1, 3, 4), (2, 4, 5))
_T10 = {}
_names = [(0, 0, 0), (0, 0, 1), (0, 0, 2), (0, 1, 1), (0, 1, 2), (0, 2, 2), (1, 1, 1), (1, 1, 2), (1, 2, 2), (2, 2, 2)]
for _i, _t in enumerate(_names):
    _T10[_t] = _i

def _workspace(nbytes, device):
    """Scratch for the partial sums: a fresh block from torch's caching allocator per call, so that concurrent
    callers on different streams never share it and a captured CUDA graph owns the block it was captured with."""
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


class _WeightedMoments(torch.autograd.Function):
    """(Wt [B,N,K], P [B,N,3], X [B,N,3]) -> M [B,K,32] float64.  Linear in Wt."""

    @staticmethod
    def forward(ctx, Wt, P, X):
        if not Wt.is_cuda:
            raise RuntimeError("CPU not supported")
        B, N, K = Wt.shape
        Wt = Wt.detach().float().contiguous()
        P = P.detach().float().contiguous()
        X_ = X.detach().float().contiguous()
        M = torch.empty(B, K, NF, dtype=torch.float64, device=Wt.device)
        L = _lib.lib()
        with torch.cuda.device(Wt.device):
            ws = _workspace(L.cpfn_moments_workspace_bytes(B, N, K), Wt.device)
            _lib.check(L.cpfn_weighted_moments(P.data_ptr(), X_.data_ptr(), Wt.data_ptr(), B, N, K, M.data_ptr(),
                                               ws.data_ptr(), ws.numel(),
                                               torch.cuda.current_stream(Wt.device).cuda_stream), "weighted_moments")
        cuda_ops.count_launches(2)
        ctx.save_for_backward(Wt, P, X_)
        return M

    @staticmethod
    def backward(ctx, dM):
        Wt, P, X = ctx.saved_tensors
        B, N, K = Wt.shape
        dM = dM.detach().double().contiguous()
        need_w, need_x = ctx.needs_input_grad[0], ctx.needs_input_grad[2]
        dW = torch.empty_like(Wt) if need_w else None
        dX = torch.empty_like(X) if need_x else None
        with torch.cuda.device(Wt.device):
            _lib.check(_lib.lib().cpfn_weighted_moments_grad(
                P.data_ptr(), X.data_ptr(), Wt.data_ptr(), dM.data_ptr(), B, N, K,
                dW.data_ptr() if need_w else None, dX.data_ptr() if need_x else None,
                torch.cuda.current_stream(Wt.device).cuda_stream), "weighted_moments_grad")
        cuda_ops.count_launches(2)
        return dW, None, dX


def weighted_moments(Wt, P, X):
    return _WeightedMoments.apply(Wt, P, X)


# ---- tiny-tensor algebra (float64, autograd) ---------------------------------------------------

def _sym3(m6):
    """[...,6] (xx,xy,xz,yy,yz,zz) -> [...,3,3]."""
    idx = torch.tensor(_SYM6, device=m6.device)
    return m6[..., idx]


def _t3(m10):
    """[...,10] -> fully symmetric [...,3,3,3]."""
    idx = torch.tensor([[[_T10[tuple(sorted((i, j, k)))] for k in range(3)] for j in range(3)] for i in range(3)],
                       device=m10.device)
    return m10[..., idx]


def sym_eigh(M, vectors=True):
    """Symmetric [*,D,D] float64 (D = 2, 3) -> (eigenvalues ascending [*,D], eigenvectors in columns [*,D,D] | None),
    the contract of torch.linalg.eigh, through cpfn_sym_eigh_small.  Not differentiable (callers detach or carry their
    own backward).  CPU tensors raise: there is no CPU path."""
    if not M.is_cuda:
        raise RuntimeError("CPU not supported")
    D = M.shape[-1]
    A = M.detach().to(torch.float64).contiguous()
    n = A.numel() // (D * D)
    lam = torch.empty(A.shape[:-1], dtype=torch.float64, device=A.device)
    Q = torch.empty_like(A) if vectors else None
    with torch.cuda.device(A.device):
        _lib.check(_lib.lib().cpfn_sym_eigh_small(A.data_ptr(), n, D, lam.data_ptr(), Q.data_ptr() if vectors else None,
                                                  torch.cuda.current_stream(A.device).cuda_stream), "sym_eigh_small")
    cuda_ops.count_launches(1)
    return lam, Q


def _solve_raw(A, b, transpose):
    D = A.shape[-1]
    A = A.detach().to(torch.float64).contiguous()
    b = b.detach().to(torch.float64).contiguous()
    x = torch.empty_like(b)
    with torch.cuda.device(A.device):
        _lib.check(_lib.lib().cpfn_small_solve(A.data_ptr(), b.data_ptr(), b.numel() // D, D, int(transpose), x.data_ptr(),
                                               torch.cuda.current_stream(A.device).cuda_stream), "small_solve")
    cuda_ops.count_launches(1)
    return x


class _SmallSolve(torch.autograd.Function):
    """x = A^-1 b for [*,D,D], [*,D] (D <= 3) through cpfn_small_solve; backward: db = A^-T g, dA = -db x^T."""

    @staticmethod
    def forward(ctx, A, b):
        if not A.is_cuda:
            raise RuntimeError("CPU not supported")
        x = _solve_raw(A, b, False)
        ctx.save_for_backward(A.detach(), x)
        return x

    @staticmethod
    def backward(ctx, g):
        A, x = ctx.saved_tensors
        gb = _solve_raw(A, g, True)
        return -gb.unsqueeze(-1) * x.unsqueeze(-2), gb


def small_solve(A, b):
    return _SmallSolve.apply(A, b)


def guard_one_over_matrix(M, min_abs_value=1e-10):
    """SPFN/differentiable_tls.py:8-17."""
    n = M.shape[-1]
    eye = torch.eye(n, dtype=M.dtype, device=M.device)
    up = torch.triu(torch.clamp(M, min=min_abs_value), diagonal=0)
    low = torch.tril(torch.clamp(M, max=-min_abs_value), diagonal=0)
    return 1.0 / (up + low + eye) - eye


class _SvdVLastColumn(torch.autograd.Function):
    """Custom_svd_v_colum (SPFN/differentiable_tls.py:123-143) for symmetric matrices [*,n,n]."""

    @staticmethod
    def forward(ctx, M):
        lam, Q = sym_eigh(M)                                # symmetric: singular values = |eigenvalues|
        s = lam.abs()
        order = torch.argsort(s, dim=-1, descending=True, stable=True)
        s = torch.gather(s, -1, order)
        lam = torch.gather(lam, -1, order)
        V = torch.gather(Q, -1, order.unsqueeze(-2).expand_as(Q))
        sgn = torch.where(lam < 0, -torch.ones_like(lam), torch.ones_like(lam))
        v = V[..., -1]
        lead = torch.gather(v, -1, v.abs().argmax(dim=-1, keepdim=True))      # deterministic sign
        flip = torch.where(lead < 0, -torch.ones_like(lead), torch.ones_like(lead))
        V = torch.cat([V[..., :-1], V[..., -1:] * flip.unsqueeze(-1)], dim=-1)
        U = V * sgn.unsqueeze(-2)
        ctx.save_for_backward(U, s, V)
        return V[..., -1].contiguous()

    @staticmethod
    def backward(ctx, grad_out):
        U, s, V = ctx.saved_tensors
        grad_v = torch.zeros_like(V)
        grad_v[..., -1] = grad_out
        s2 = s ** 2
        K = guard_one_over_matrix(s2.unsqueeze(-1) - s2.unsqueeze(-2))
        inner = K.transpose(-1, -2) * (V.transpose(-1, -2) @ grad_v)
        inner = (inner + inner.transpose(-1, -2)) / 2
        # gradients through s and u are ignored by design (differentiable_tls.py:141)
        return U @ (2 * torch.diag_embed(s) @ inner) @ V.transpose(-1, -2)


def svd_v_last_column(M):
    return _SvdVLastColumn.apply(M)


def guarded_solve(AtA, Atb, condition_number_cap=1e5, ls_l2_regularizer=1e-8):
    """Normal-equation form of guarded_matrix_solve_ls (SPFN/geometry_utils.py:131-141).
    AtA [*,D,D], Atb [*,D] -> x [*,D]."""
    D = AtA.shape[-1]
    s = sym_eigh(AtA, vectors=False)[0].abs()
    mask = ((s.max(dim=-1)[0] / s.min(dim=-1)[0]) < condition_number_cap).to(AtA.dtype)
    eye = torch.eye(D, dtype=AtA.dtype, device=AtA.device)
    A = AtA * mask[..., None, None] + ls_l2_regularizer * eye
    b = Atb * mask[..., None]
    return small_solve(A, b)


def compute_consistent_plane_frame(normal):
    """SPFN/geometry_utils.py:8-27 (any floating dtype)."""
    eye = torch.eye(3, dtype=normal.dtype, device=normal.device)
    y_axes = torch.stack([torch.cross(normal, eye[q].expand_as(normal), dim=-1) for q in range(3)], dim=0)
    chosen = torch.argmax(torch.norm(y_axes, dim=-1), dim=0)
    y = torch.gather(y_axes, 0, chosen.view(1, *chosen.shape, 1).expand(1, *normal.shape)).squeeze(0)
    y = torch.nn.functional.normalize(y, p=2, dim=-1, eps=1e-12)
    return torch.cross(y, normal, dim=-1), y


def _sphere_from_moments(R1, R2, R3, Sw, R1c, R2c, R3c, Swc, eps=1e-10):
    """weighted_sphere_fitting (SPFN/geometry_utils.py:209-223) for D = 2 or 3 from raw moments:
    R1 [*,D] = sum w p, R2 [*,D,D] = sum w p p^T, Sw = sum w; the *c versions use w' = max(w, 1e-10)
    (the LS row weights); R3c [*,D,D,D] = sum w' p p p."""
    denom = torch.clamp(Sw, min=eps)
    mu = R1 / denom[..., None]
    m2 = torch.diagonal(R2, dim1=-2, dim2=-1).sum(-1) / denom
    # A = 2 (mu - p), b = m2 - |p|^2, rows weighted by w'
    Sp = R2c - mu[..., :, None] * R1c[..., None, :] - R1c[..., :, None] * mu[..., None, :] \
        + Swc[..., None, None] * mu[..., :, None] * mu[..., None, :]
    s1p = R1c - mu * Swc[..., None]                                   # sum w' (p - mu)
    trc = torch.diagonal(R2c, dim1=-2, dim2=-1).sum(-1)               # sum w' |p|^2
    t3 = torch.diagonal(R3c, dim1=-2, dim2=-1).sum(-1)                # sum w' p |p|^2
    Atb = -2 * m2[..., None] * s1p + 2 * (t3 - mu * trc[..., None])
    center = guarded_solve(4 * Sp, Atb)
    r2 = (torch.diagonal(R2, dim1=-2, dim2=-1).sum(-1) - 2 * (center * R1).sum(-1) + (center * center).sum(-1) * Sw) / denom
    return center, r2


def compute_parameters(P, W, X, classes, eps=1e-10):
    """Differentiable fitters.  P [B,N,3], W [B,N,K], X [B,N,3] -> dict (float32 outputs)."""
    Wc = torch.clamp(W, min=eps)                                      # sqrt_W^2 of guarded_matrix_solve_ls
    M = weighted_moments(W, P, X)
    Mc = weighted_moments(Wc, P, X)
    Sw, Swc = M[..., 0], Mc[..., 0]
    denom = torch.clamp(Sw, min=eps)
    R1, R2, R1c, R2c = M[..., _P1:_P1 + 3], _sym3(M[..., _P2:_P2 + 6]), Mc[..., _P1:_P1 + 3], _sym3(Mc[..., _P2:_P2 + 6])
    R3c = _t3(Mc[..., _P3:_P3 + 10])
    mu = R1 / denom[..., None]
    out = {}
    f32 = lambda t: t.to(torch.float32)
    if "plane" in classes:
        S = R2 - mu[..., :, None] * R1[..., None, :] - R1[..., :, None] * mu[..., None, :] \
            + Sw[..., None, None] * mu[..., :, None] * mu[..., None, :]
        n = svd_v_last_column(S)
        out["plane_normal"], out["plane_center"] = f32(n), f32((n * mu).sum(-1))
    if "sphere" in classes:
        c, r2 = _sphere_from_moments(R1, R2, None, Sw, R1c, R2c, R3c, Swc)
        out["sphere_center"], out["sphere_radius_squared"] = f32(c), f32(r2)
    Mxx = _sym3(M[..., _X2:_X2 + 6])
    if "cylinder" in classes:
        n = svd_v_last_column(Mxx)                                    # TLS on the normals, uncentred
        xa, ya = compute_consistent_plane_frame(n)
        F = torch.stack([xa, ya], dim=-1)                             # [B,K,3,2]
        Ft = F.transpose(-1, -2)
        proj1 = lambda v: (Ft @ v.unsqueeze(-1)).squeeze(-1)
        proj2 = lambda m: Ft @ m @ F
        R3q = torch.einsum("...ia,...jb,...lc,...ijl->...abc", F, F, F, R3c)
        cq, r2 = _sphere_from_moments(proj1(R1), proj2(R2), None, Sw, proj1(R1c), proj2(R2c), R3q, Swc)
        center = cq[..., 0:1] * xa + cq[..., 1:2] * ya
        out["cylinder_axis"], out["cylinder_center"], out["cylinder_radius_squared"] = f32(n), f32(center), f32(r2)
    if "cone" in classes:
        AtA = _sym3(Mc[..., _X2:_X2 + 6])
        Atb = Mc[..., _XP:_XP + 3]
        apex = guarded_solve(AtA, Atb)
        mux = M[..., _X1:_X1 + 3] / denom[..., None]
        Rx1 = M[..., _X1:_X1 + 3]
        Cx = Mxx - mux[..., :, None] * Rx1[..., None, :] - Rx1[..., :, None] * mux[..., None, :] \
            + Sw[..., None, None] * mux[..., :, None] * mux[..., None, :]
        axis = svd_v_last_column(Cx)
        # cone_fitter.py:24-35 on [B,N,K]: element-wise torch (fp32 like the reference)
        apex32, axis32 = f32(apex), f32(axis)
        d = P.unsqueeze(2) - apex32.unsqueeze(1)                      # [B,N,K,3]
        dn = torch.nn.functional.normalize(d, p=2, dim=3, eps=1e-12)
        dot = torch.sum(axis32.unsqueeze(1) * dn, dim=3)              # [B,N,K]
        sgn = torch.sign(torch.sum(W * dot, dim=1))
        sgn = sgn + (sgn == 0.0).float()
        axis32 = axis32 * sgn.unsqueeze(2)
        ang = torch.acos(torch.clamp(torch.abs(dot), min=-1.0 + 1e-6, max=1.0 - 1e-6))
        half = torch.sum(W * ang, dim=1) / (torch.sum(W, dim=1) + eps)
        half = torch.clamp(half, min=1e-3, max=math.pi / 2 - 1e-3)
        out["cone_apex"], out["cone_axis"], out["cone_half_angle"] = apex32, axis32, half
    return out
