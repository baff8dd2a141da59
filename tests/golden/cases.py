"""Seeded inputs shared by the golden-vector generators and the parity tests.

Inputs are regenerated from seeds (cpfn_b200.synth, numpy default_rng), only the
reference OUTPUTS are stored in the .npz fixtures.
"""
import numpy as np

from cpfn_b200 import synth


def fps_cases():
    """name -> (xyz [B,N,3], nsamples)."""
    P_shape = synth.shape_batch(2, 8192, seed=11)[0]
    return {
        "uniform_8192_512": (synth.uniform_cloud(2, 8192, seed=1), 512),
        "shape_8192_512": (P_shape, 512),
        "lattice_1000_300": (synth.lattice_cloud(3, 1000, seed=2), 300),
        "lattice_259_100": (synth.lattice_cloud(2, 259, seed=3), 100),
        "uniform_512_128": (synth.uniform_cloud(4, 512, seed=4), 128),
        "lattice_40_40": (synth.lattice_cloud(2, 40, seed=5, pitch=4), 40),
        "uniform_12000_64": (synth.uniform_cloud(1, 12000, seed=6), 64),
        "uniform_20000_48": (synth.uniform_cloud(1, 20000, seed=7), 48),
    }


def _take(xyz, idx):
    return np.take_along_axis(xyz, idx[..., None].astype(np.int64), 1)


def ball_cases(fps):
    """name -> (new_xyz, xyz, radius, nsample); `fps(xyz, m)` supplies centroids."""
    u = synth.uniform_cloud(2, 8192, seed=1)
    s = synth.shape_batch(2, 8192, seed=11)[0]
    lat = synth.lattice_cloud(2, 1000, seed=2)
    c1 = _take(s, fps(s, 512))
    cases = {
        "sa1_shape": (c1, s, 0.2, 64),
        "sa2_shape": (_take(c1, fps(c1, 128)), c1, 0.4, 64),
        "sa1_uniform": (_take(u, fps(u, 512))[:, :128], u, 0.2, 64),
        "lattice_ties": (lat[:, :200], lat, 0.25, 32),
        "tiny_radius": (u[:, :64] + np.float32(0.5), u, 1e-3, 16),
        "sparse_hits": (u[:, :64], u[:, :300], 0.2, 48),
    }
    return cases


def three_nn_cases(fps):
    s = synth.shape_batch(1, 4096, seed=12)[0]
    lat = synth.lattice_cloud(1, 1000, seed=8, pitch=6)
    return {
        "fp3_like": (s, _take(s, fps(s, 512))),
        "lattice_ties": (lat, lat[:, :100]),
        "two_known": (s[:, :50], s[:, :2]),
    }


def interp_inputs():
    rng = np.random.default_rng(21)
    B, C, M, n = 2, 16, 128, 1024
    pts = rng.normal(size=(B, C, M)).astype(np.float32)
    idx = rng.integers(0, M, size=(B, n, 3)).astype(np.int32)
    w = rng.random(size=(B, n, 3)).astype(np.float32)
    w /= w.sum(2, keepdims=True)
    g = rng.normal(size=(B, C, n)).astype(np.float32)
    gidx = rng.integers(0, M, size=(B, 32, 8)).astype(np.int32)
    gg = rng.normal(size=(B, C, 32, 8)).astype(np.float32)
    return dict(pts=pts, idx=idx, w=w, g=g, gidx=gidx, gg=gg, M=M)


def fitter_cases():
    """name -> (P [B,N,3], W [B,N,K], X [B,N,3]).  `selfcheck` re-uses the seeds and
    distributions of the reference's inline self-checks (SPFN/plane_fitter.py:30-39,
    cylinder_fitter.py:51-63: np.random.seed(0), randn points, rand weights,
    normalised randn normals) at a smaller batch."""
    out = {}
    rs = np.random.RandomState(0)
    P = rs.randn(4, 1024, 3)
    W = rs.rand(4, 1024, 12)
    X = rs.randn(4, 1024, 3)
    X = X / np.linalg.norm(X, axis=2, keepdims=True)
    out["selfcheck"] = (P.astype(np.float32), W.astype(np.float32), X.astype(np.float32))
    P, X, W, _ = synth.shape_batch(2, 2048, seed=31, k_slots=24)
    out["shape_2048_k24"] = (P, W, X)
    P, X, W, _ = synth.shape_batch(1, 8192, seed=1234, k_slots=24)   # BASELINE config 1
    out["config1_8192_k24"] = (P, W, X)
    P, X, W, I = synth.shape_batch(2, 4096, seed=33, k_slots=28)
    hard = np.zeros_like(W)
    np.put_along_axis(hard, W.argmax(2)[..., None], 1.0, axis=2)      # one-hot, slots >= 12 empty
    out["onehot_4096_k28"] = (P, hard, X)
    return out


# synth.shape_cloud: instances 0-2 planes, 3-5 spheres, 6-8 cylinders, 9-11 cones.
_TYPE_SLOTS = {"plane": (0, 3), "sphere": (3, 6), "cylinder": (6, 9), "cone": (9, 12)}


def fit_mask(case, key, W):
    """[B,K] mask of (cloud, slot) pairs on which parameter `key` is well-posed.

    Fitting a cylinder or a cone to a plane patch (all normals equal) or a plane to a
    sphere is rank-deficient: the reference returns an arbitrary member of a null
    space there, so parity is asserted on the slots whose ground-truth primitive has
    the type being fitted (every slot for the unstructured `selfcheck` case), and
    never on slots without support."""
    B, _, K = W.shape
    live = W.sum(1) > 1.0
    if case == "selfcheck":
        return live
    lo, hi = _TYPE_SLOTS[key.split("_")[0]]
    m = np.zeros((B, K), dtype=bool)
    m[:, lo:hi] = True
    return m & live


network_state = synth.network_state      # deterministic PointNet2 parameters (shared with bench.py)


def network_input(batch=2, n_points=1024, seed=51):
    return synth.shape_batch(batch, n_points, seed=seed)[0]


def grad_case():
    """(P, W, X) for the fitter-gradient parity: every slot is a live primitive."""
    P, X, W, _ = synth.shape_batch(2, 1024, seed=77, k_slots=12)
    return P, W, X


def fitter_loss(params, W_np, torch):
    """Scalar, sign-invariant loss over the well-posed slots of grad_case() (torch tensors in `params`):
    linear in the sign-determinate parameters, quadratic (n n^T, c^2) in the sign-ambiguous ones."""
    rng = np.random.default_rng(5)
    total = 0.0
    for key in sorted(params):
        v = params[key]
        mask = torch.as_tensor(fit_mask("grad", key, W_np).astype(np.float32), device=v.device).to(v.dtype)
        if key in ("plane_normal", "cylinder_axis"):
            G = torch.as_tensor(rng.normal(size=(3, 3)).astype(np.float32), device=v.device).to(v.dtype)
            total = total + (mask * torch.einsum("bki,ij,bkj->bk", v, G, v)).sum()
        elif key == "plane_center":
            g = float(rng.normal())
            total = total + (mask * g * v * v).sum()
        else:
            G = torch.as_tensor(rng.normal(size=tuple(v.shape)).astype(np.float32), device=v.device).to(v.dtype)
            m = mask if v.dim() == 2 else mask.unsqueeze(-1)
            total = total + (m * G * v).sum()
    return total


def patch_cases():
    """name -> (points_lr [n_lr,3], points_hr [N,3], pool_indices, pool_labels, k, max_patches, np.random seed).
    Low-res = a strided subset of the high-res shape cloud (as the reference's low-res files are a
    sub-sampling of the high-res ones); fp32 distances of distinct points do tie now and then, see assert_patches_equivalent."""
    out = {}
    for name, N, n_lr, k, mp, seed in (("small", 4096, 512, 256, 8, 3), ("odd_k", 5000, 400, 300, 6, 4),
                                       ("mid", 20000, 1024, 2048, 40, 5)):
        hr = synth.shape_cloud(N, 900 + seed)[0].astype(np.float32)
        lr = np.ascontiguousarray(hr[:: N // n_lr][:n_lr])
        rng = np.random.RandomState(seed)
        pool = np.sort(rng.choice(n_lr, size=n_lr // 2, replace=False))
        labels = rng.randint(0, 5, size=len(pool))
        out[name] = (lr, hr, pool, labels, k, mp, 100 + seed)
    return out


def assert_patches_equivalent(got, ref, hr):
    """Patch index rows are equal up to the order of EQUAL distances (the reference's introsort does not
    define it): same shape, every row holds the same points, and position by position the distance to the
    row's seed (= its first, zero-distance entry: the low-res points are a subset of the high-res ones) is
    bit-identical."""
    assert got.shape == ref.shape, (got.shape, ref.shape)
    for r in range(ref.shape[0]):
        if np.array_equal(got[r], ref[r]):
            continue
        assert got[r, 0] == ref[r, 0]
        seed = hr[ref[r, 0]]
        dg = np.linalg.norm(seed[None] - hr[got[r]], axis=1)
        dr = np.linalg.norm(seed[None] - hr[ref[r]], axis=1)
        assert np.array_equal(dg, dr), r
        assert np.array_equal(np.sort(got[r]), np.sort(ref[r])), r


def merging_cases():
    """name -> dict(W [nb,Np,Kl] soft-maxed memberships, X [nb,Np,3], T [nb,Np,4], idx int64 [nb,Np] (unique per
    patch), S int64 one-hot [Ng,Kg] object labels, obj_normals, obj_types).  Patches are spatial neighbourhoods of
    a shape cloud whose memberships follow the cloud's ground-truth primitives (+ noise), so patches overlap,
    share primitives and leave part of the cloud uncovered -- the situation the merge is built for."""
    out = {}
    for name, Ng, nb, Np, Kl, Kg, seed in (("small", 3000, 5, 512, 6, 7, 11), ("wide", 6000, 9, 700, 21, 28, 12),
                                           ("one_patch", 1500, 1, 256, 4, 5, 13)):
        rng = np.random.RandomState(seed)
        P, Xn, _, I = synth.shape_batch(1, Ng, seed=seed)
        P, Xn, I = P[0], Xn[0], I[0]
        idx = np.empty((nb, Np), np.int64)
        W = np.empty((nb, Np, Kl), np.float32)
        for b in range(nb):
            c = P[rng.randint(Ng)]
            idx[b] = np.argsort(np.linalg.norm(P - c[None], axis=1), kind="stable")[:Np]
            local = I[idx[b]] % Kl                                   # local slot of the point's primitive
            logits = 6.0 * np.eye(Kl, dtype=np.float32)[local] + rng.randn(Np, Kl).astype(np.float32)
            e = np.exp(logits - logits.max(1, keepdims=True))
            W[b] = e / e.sum(1, keepdims=True)
        Xp = Xn[idx] + 0.05 * rng.randn(nb, Np, 3).astype(np.float32)
        Xp /= np.linalg.norm(Xp, axis=2, keepdims=True)
        T = rng.randn(nb, Np, 4).astype(np.float32)
        S = np.eye(Kg, dtype=np.int64)[I % Kg]
        out[name] = dict(W=W, X=Xp.astype(np.float32), T=T, idx=idx, S=S, obj_normals=Xn.astype(np.float32),
                         obj_types=rng.randn(Ng, 4).astype(np.float32))
    return out


def residue_case(seed=21, B=2, K=6, Kp=6, n_pts=96, N=700):
    """Seeded inputs of compute_residue_loss / compute_P_coverage: plausible primitive parameters (unit normals and
    axes, positive radii, half angles inside the clamp), a permutation-like matching with repeats, points around the
    unit ball, one primitive with a point exactly on the cone apex (the F.normalize eps branch)."""
    rng = np.random.RandomState(seed)
    unit = lambda a: (a / np.linalg.norm(a, axis=-1, keepdims=True)).astype(np.float32)
    params = {
        "plane_normal": unit(rng.randn(B, Kp, 3)), "plane_center": rng.randn(B, Kp).astype(np.float32) * 0.3,
        "sphere_center": rng.randn(B, Kp, 3).astype(np.float32) * 0.3,
        "sphere_radius_squared": (rng.rand(B, Kp).astype(np.float32) * 0.5 + 0.01),
        "cylinder_axis": unit(rng.randn(B, Kp, 3)), "cylinder_center": rng.randn(B, Kp, 3).astype(np.float32) * 0.3,
        "cylinder_radius_squared": (rng.rand(B, Kp).astype(np.float32) * 0.3 + 0.01),
        "cone_apex": rng.randn(B, Kp, 3).astype(np.float32) * 0.5, "cone_axis": unit(rng.randn(B, Kp, 3)),
        "cone_half_angle": (rng.rand(B, Kp).astype(np.float32) * 1.2 + 0.1),
    }
    params["sphere_radius_squared"][0, 0] = -0.04              # sqrt_safe takes |x|
    matching = rng.randint(0, Kp, (B, K)).astype(np.int64)
    points = (rng.randn(B, K, n_pts, 3) * 0.5).astype(np.float32)
    points[0, 1, 0] = params["cone_apex"][0, matching[0, 1]]   # v = 0
    T_gt = rng.randint(0, 4, (B, K)).astype(np.int64)
    P = (rng.randn(B, N, 3) * 0.5).astype(np.float32)
    return params, matching, points, T_gt, P


def lowres_cases():
    """name -> (points float32 [N,3], labels int32 [N], nb_query_points, np.random seed).  Shape clouds with their
    ground-truth primitive labels; 'lattice' has many exactly equal distances (first-maximum tie-break) and
    duplicated points."""
    out = {}
    for name, N, m, seed in (("small", 3000, 200, 31), ("mid", 20000, 512, 32)):
        P, _, _, I = synth.shape_batch(1, N, seed=seed)
        out[name] = (P[0].astype(np.float32), I[0].astype(np.int32), m, 200 + seed)
    lat = synth.lattice_cloud(1, 2048, seed=9)[0].astype(np.float32)
    lat = np.concatenate([lat, lat[:100]])
    out["lattice"] = (lat, (np.arange(len(lat)) % 7).astype(np.int32), 150, 233)
    return out


def seg_cases():
    """name -> (W_pred float32 [B,N,K], I_gt int64 [B,N]) for hungarian_matching / compute_miou_loss.
    'soft': soft-maxed memberships of shape clouds (generic costs, no ties); 'hard': one-hot predictions that are a
    permutation of the ground truth plus unused slots and an absent ground-truth label (all-zero cost rows and exact
    ties: the solver's tie-breaking decides); 'background': points labelled -1; 'few': two ground-truth labels."""
    out = {}
    P, X, W, I = synth.shape_batch(3, 2048, seed=61, k_slots=28)
    out["soft"] = (W.astype(np.float32), I.astype(np.int64))
    rng = np.random.RandomState(62)
    B, N, K = 4, 1500, 12
    I = rng.randint(0, 9, size=(B, N)).astype(np.int64)
    I[I == 4] = 5                                                  # label 4 never occurs: an all-zero cost row
    perm = np.stack([rng.permutation(K) for _ in range(B)])
    hard = np.zeros((B, N, K), np.float32)
    for b in range(B):
        hard[b, np.arange(N), perm[b][I[b]]] = 1.0
    out["hard"] = (hard, I)
    W2 = rng.rand(2, 1000, 24).astype(np.float32)
    W2 /= W2.sum(2, keepdims=True)
    I2 = rng.randint(-1, 10, size=(2, 1000)).astype(np.int64)
    out["background"] = (W2, I2)
    W3 = rng.rand(5, 700, 28).astype(np.float32)
    I3 = rng.randint(0, 2, size=(5, 700)).astype(np.int64)
    out["few"] = (W3, I3)
    return out
