"""numpy fp32 restatement of SPFN's weighted total-least-squares fitters.

TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Every function follows the reference
line by line (paths relative to the reference tree) with numpy float32 arrays,
``np.linalg.svd`` / ``np.linalg.solve`` standing in for ``torch.svd`` /
``torch.solve`` (both are LAPACK on CPU).

Parity pin: tests/golden/ref_fitters.npz, produced by importing the UNMODIFIED
reference ``SPFN`` package in the dev container
(tests/golden/make_ref_fitters_golden.py, with the two torch-2.11 shims the
reference needs: ``torch.solve`` and a CPU-safe ``Tensor.get_device``).
Eigen-vector outputs are sign-ambiguous: compare up to sign.
"""
import numpy as np

F = np.float32


def solve_weighted_tls(A, W):
    """SPFN/differentiable_tls.py:200-209.  A [B,N,3], W [B,N] -> x [B,3]:
    last right-singular vector of M = sum_n w a a^T."""
    A = A.astype(F)
    W = W.astype(F)
    A_p = A[:, :, None, :] * A[:, :, :, None]           # BxNx3x3
    M = np.sum(W[:, :, None, None] * A_p, axis=1, dtype=F)  # Bx3x3
    _, _, vt = np.linalg.svd(M)                          # Custom_svd_v_colum :123-130
    return vt[:, -1, :].astype(F)


def weighted_plane_fitting(P, W, division_eps=1e-10):
    """SPFN/geometry_utils.py:74-84."""
    P = P.astype(F)
    W = W.astype(F)
    WP = P * W[:, :, None]
    W_sum = np.sum(W, axis=1, keepdims=True, dtype=F)
    mean = np.sum(WP, axis=1, dtype=F) / np.maximum(W_sum, F(division_eps))
    A = P - mean[:, None, :]
    n = solve_weighted_tls(A, W)
    c = np.sum(n * mean, axis=1, dtype=F)
    return n, c


def guarded_matrix_solve_ls(A, b, W, condition_number_cap=1e5, sqrt_eps=1e-10,
                            ls_l2_regularizer=1e-8):
    """SPFN/geometry_utils.py:121-142.  A [B,N,D], b [B,N,1], W [B,N] -> x [B,D]."""
    A = A.astype(F)
    b = b.astype(F)
    dim = A.shape[2]
    sqrt_W = np.sqrt(np.maximum(W.astype(F), F(sqrt_eps)))[:, :, None]
    A = A * sqrt_W
    b = b * sqrt_W
    AtA = np.matmul(A.transpose(0, 2, 1), A).astype(F)
    s = np.linalg.svd(AtA, compute_uv=False)
    with np.errstate(divide="ignore", invalid="ignore"):
        mask = (s[:, 0] / s[:, -1] < condition_number_cap).astype(F)
    AtA = AtA * mask[:, None, None] + F(ls_l2_regularizer) * np.eye(dim, dtype=F)[None]
    Atb = np.matmul(A.transpose(0, 2, 1) * mask[:, None, None], b).astype(F)
    x = np.linalg.solve(AtA, Atb)
    return x[:, :, 0].astype(F)


def weighted_sphere_fitting(P, W, division_eps=1e-10):
    """SPFN/geometry_utils.py:209-223.  P [B,N,D] (D = 2 or 3), W [B,N]."""
    P = P.astype(F)
    W = W.astype(F)
    W_sum = np.sum(W, axis=1, dtype=F)
    P_sqr = np.sum(P ** 2, axis=2, dtype=F)
    WP_sqr_sum = np.sum(W * P_sqr, axis=1, dtype=F)
    denom = np.maximum(W_sum, F(division_eps))
    b = ((WP_sqr_sum / denom)[:, None] - P_sqr)[:, :, None]
    WP_sum = np.sum(W[:, :, None] * P, axis=1, dtype=F)
    A = 2 * ((WP_sum / denom[:, None])[:, None, :] - P)
    center = guarded_matrix_solve_ls(A, b, W)
    d = P - center[:, None, :]
    r_sqr = np.sum(W * np.sum(d ** 2, axis=2, dtype=F), axis=1, dtype=F) / denom
    return center, r_sqr.astype(F)


def compute_consistent_plane_frame(normal):
    """SPFN/geometry_utils.py:8-27.  normal [B,3] -> x_axes, y_axes [B,3]."""
    normal = normal.astype(F)
    cands = np.eye(3, dtype=F)
    y_axes = np.stack([np.cross(normal, np.broadcast_to(c, normal.shape)) for c in cands], axis=0)
    norms = np.linalg.norm(y_axes, axis=2)
    chosen = np.argmax(norms, axis=0)          # first maximum on ties, like torch.argmax
    y = y_axes[chosen, np.arange(normal.shape[0])]
    y = y / np.maximum(np.linalg.norm(y, axis=1, keepdims=True), F(1e-12))
    x = np.cross(y, normal)
    return x.astype(F), y.astype(F)


def _tile(P, W):
    B, N, _ = P.shape
    K = W.shape[2]
    W_r = np.ascontiguousarray(W.transpose(0, 2, 1)).reshape(B * K, N)
    P_t = np.broadcast_to(P[:, None], (B, K, N, P.shape[2])).reshape(B * K, N, P.shape[2])
    return P_t, W_r, B, K


def plane_compute_parameters(P, W):
    """SPFN/plane_fitter.py:9-17 -> n [B,K,3], c [B,K]."""
    P_t, W_r, B, K = _tile(P, W)
    n, c = weighted_plane_fitting(P_t, W_r)
    return n.reshape(B, K, 3), c.reshape(B, K)


def sphere_compute_parameters(P, W):
    """SPFN/sphere_fitter.py:9-19 -> center [B,K,3], radius_squared [B,K]."""
    P_t, W_r, B, K = _tile(P, W)
    c, r2 = weighted_sphere_fitting(P_t, W_r)
    return c.reshape(B, K, 3), r2.reshape(B, K)


def cylinder_compute_parameters(P, W, X):
    """SPFN/cylinder_fitter.py:10-28 -> axis [B,K,3], center [B,K,3], radius_squared [B,K]."""
    X_t, W_r, B, K = _tile(X, W)
    n = solve_weighted_tls(X_t, W_r).reshape(B, K, 3)
    x_axes, y_axes = compute_consistent_plane_frame(n.reshape(B * K, 3))
    x_axes = x_axes.reshape(B, K, 3)
    y_axes = y_axes.reshape(B, K, 3)
    P = P.astype(F)
    x_coord = np.sum(P[:, None] * x_axes[:, :, None], axis=3, dtype=F)
    y_coord = np.sum(P[:, None] * y_axes[:, :, None], axis=3, dtype=F)
    P_proj = np.stack([x_coord, y_coord], axis=3).reshape(B * K, P.shape[1], 2)
    cc, r2 = weighted_sphere_fitting(P_proj, W_r)
    cc = cc.reshape(B, K, 2)
    center = cc[:, :, 0:1] * x_axes + cc[:, :, 1:2] * y_axes
    return n, center.astype(F), r2.reshape(B, K)


def cone_compute_parameters(P, W, X, div_eps=1e-10):
    """SPFN/cone_fitter.py:12-36 -> apex [B,K,3], axis [B,K,3], half_angle [B,K]."""
    P = P.astype(F)
    X = X.astype(F)
    W = W.astype(F)
    A, W_r, B, K = _tile(X, W)
    N = P.shape[1]
    b = np.broadcast_to(np.sum(P * X, axis=2, dtype=F)[:, None], (B, K, N)).reshape(B * K, N, 1)
    apex = guarded_matrix_solve_ls(A, b, W_r).reshape(B, K, 3)
    plane_n, _ = weighted_plane_fitting(A, W_r)
    axis = plane_n.reshape(B, K, 3)
    d = P[:, :, None, :] - apex[:, None, :, :]                       # BxNxKx3
    dn = d / np.maximum(np.linalg.norm(d, axis=3, keepdims=True), F(1e-12))
    dot = np.sum(axis[:, None] * dn, axis=3, dtype=F)                # BxNxK
    sgn = np.sign(np.sum(W * dot, axis=1, dtype=F))
    sgn = sgn + (sgn == 0).astype(F)
    axis = axis * sgn[:, :, None]
    ang = np.arccos(np.clip(np.abs(dot), F(-1.0 + 1e-6), F(1.0 - 1e-6)))
    half = np.sum(W * ang, axis=1, dtype=F) / (np.sum(W, axis=1, dtype=F) + F(div_eps))
    half = np.clip(half, F(1e-3), F(np.pi / 2 - 1e-3))
    return apex, axis.astype(F), half.astype(F)


def compute_parameters(P, W, X, classes=("plane", "sphere", "cylinder", "cone")):
    """SPFN/losses_implementation.py:255-278 (same dictionary keys)."""
    out = {}
    for c in classes:
        if c == "plane":
            out["plane_normal"], out["plane_center"] = plane_compute_parameters(P, W)
        elif c == "sphere":
            out["sphere_center"], out["sphere_radius_squared"] = sphere_compute_parameters(P, W)
        elif c == "cylinder":
            (out["cylinder_axis"], out["cylinder_center"],
             out["cylinder_radius_squared"]) = cylinder_compute_parameters(P, W, X)
        elif c == "cone":
            out["cone_apex"], out["cone_axis"], out["cone_half_angle"] = cone_compute_parameters(P, W, X)
        else:
            raise NotImplementedError
    return out


def compute_parameters_f64(P, W, X, classes=("plane", "sphere", "cylinder", "cone")):
    """The same algebra carried out in float64: the conditioning yardstick of the parity tests.  The distance of
    the float32 restatement above (= the reference's own arithmetic) from this result says how much of a difference
    on a given slot is rounding noise amplified by an ill-conditioned fit (a cylinder axis from near-parallel
    normals ...) rather than a disagreement about the algorithm."""
    global F
    saved, F = F, np.float64
    try:
        return compute_parameters(np.asarray(P, np.float64), np.asarray(W, np.float64), np.asarray(X, np.float64), classes)
    finally:
        F = saved


# --- point-to-primitive residuals (compute_residue_single of each fitter) ---------------------

def _sqrt_safe(x):
    return np.sqrt(np.abs(x) + F(1e-10))


def _acos_safe(x):
    return np.arccos(np.clip(x, F(-1.0 + 1e-6), F(1.0 - 1e-6)))


def plane_residue(n, c, p):
    """SPFN/plane_fitter.py:54-55."""
    return (np.sum(p * n, axis=-1) - c) ** 2


def sphere_residue(center, radius_squared, p):
    """SPFN/sphere_fitter.py:58-62."""
    return (_sqrt_safe(np.sum((p - center) ** 2, axis=-1)) - _sqrt_safe(radius_squared)) ** 2


def cylinder_residue(axis, center, radius_squared, p):
    """SPFN/cylinder_fitter.py:82-89."""
    d = p - center
    return (_sqrt_safe(np.sum(d ** 2, axis=-1) - np.sum(d * axis, axis=-1) ** 2)
            - _sqrt_safe(radius_squared)) ** 2


def cone_residue(apex, axis, half_angle, p):
    """SPFN/cone_fitter.py:98-103."""
    v = p - apex
    vn = v / np.maximum(np.linalg.norm(v, axis=-1, keepdims=True), F(1e-12))
    alpha = _acos_safe(np.sum(vn * axis, axis=-1))
    return np.sin(np.minimum(np.abs(alpha - half_angle), F(np.pi / 2))) ** 2 * np.sum(v * v, axis=-1)
