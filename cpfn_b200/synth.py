"""Deterministic synthetic inputs for the CPFN hot path.

No dataset ships with the reference, so parity tests and ``bench.py`` use
clouds generated here.  Everything is numpy (``default_rng(seed)``), float32,
and normalised to the unit ball the way the reference normalises its clouds
(``Utils/dataset_utils.py:26-27``), which is what the SA radii 0.2 / 0.4 assume.
"""
import numpy as np


def _unit(v):
    return v / np.maximum(np.linalg.norm(v, axis=-1, keepdims=True), 1e-12)


def _frame(rng):
    a = _unit(rng.normal(size=3))
    t = _unit(np.cross(a, rng.normal(size=3)))
    return a, t, np.cross(a, t)


def shape_cloud(n_points, seed, k_slots=28, noise=0.005, logit_gain=8.0):
    """One cloud made of 12 analytic primitives (3 planes, 3 spheres, 3 cylinders,
    3 cones, random pose).

    Returns P [N,3], X [N,3] (analytic unit normals), I [N] (instance id 0..11),
    T [12] (type id: 0 plane, 1 sphere, 2 cylinder, 3 cone), W [N,k_slots]
    (soft memberships: softmax(gain * onehot + N(0,1)), slots >= 12 stay small).
    """
    rng = np.random.default_rng(seed)
    n_inst = 12
    counts = np.full(n_inst, n_points // n_inst)
    counts[: n_points - counts.sum()] += 1
    P, X, I = [], [], []
    types = np.repeat(np.arange(4), 3)
    for inst in range(n_inst):
        m = counts[inst]
        c = rng.uniform(-0.5, 0.5, size=3)
        a, t, s = _frame(rng)
        typ = types[inst]
        if typ == 0:  # plane patch
            u, v = rng.uniform(-0.35, 0.35, size=(2, m))
            p = c + u[:, None] * t + v[:, None] * s
            nrm = np.broadcast_to(a, p.shape)
        elif typ == 1:  # sphere (cap covering most of it)
            r = rng.uniform(0.12, 0.3)
            d = _unit(rng.normal(size=(m, 3)))
            p = c + r * d
            nrm = d
        elif typ == 2:  # cylinder
            r = rng.uniform(0.08, 0.2)
            th = rng.uniform(0, 2 * np.pi, size=m)
            h = rng.uniform(-0.35, 0.35, size=m)
            d = np.cos(th)[:, None] * t + np.sin(th)[:, None] * s
            p = c + r * d + h[:, None] * a
            nrm = d
        else:  # cone, apex at c, axis a
            half = rng.uniform(0.25, 0.6)
            th = rng.uniform(0, 2 * np.pi, size=m)
            h = rng.uniform(0.08, 0.45, size=m)
            d = np.cos(th)[:, None] * t + np.sin(th)[:, None] * s
            p = c + h[:, None] * a + (h * np.tan(half))[:, None] * d
            nrm = np.cos(half) * d - np.sin(half) * a
        P.append(p)
        X.append(nrm)
        I.append(np.full(m, inst))
    P = np.concatenate(P) + rng.normal(scale=noise, size=(n_points, 3))
    X = _unit(np.concatenate(X))
    I = np.concatenate(I)
    perm = rng.permutation(n_points)
    P, X, I = P[perm], X[perm], I[perm]
    P = P - P.mean(axis=0, keepdims=True)
    P = P / np.max(np.linalg.norm(P, axis=1))
    logits = rng.normal(size=(n_points, k_slots))
    logits[np.arange(n_points), I] += logit_gain
    logits -= logits.max(axis=1, keepdims=True)
    W = np.exp(logits)
    W /= W.sum(axis=1, keepdims=True)
    return (P.astype(np.float32), X.astype(np.float32), I.astype(np.int64), types.astype(np.int64),
            W.astype(np.float32))


def shape_batch(batch, n_points, seed, k_slots=28):
    """Batch of shape clouds: P [B,N,3], X [B,N,3], W [B,N,K], I [B,N]."""
    out = [shape_cloud(n_points, seed * 1000 + b, k_slots) for b in range(batch)]
    P = np.stack([o[0] for o in out])
    X = np.stack([o[1] for o in out])
    I = np.stack([o[2] for o in out])
    W = np.stack([o[4] for o in out])
    return P, X, W, I


def uniform_cloud(batch, n_points, seed):
    """U(-1,1)^3 rescaled into the unit ball (timing / stress)."""
    rng = np.random.default_rng(seed)
    P = rng.uniform(-1.0, 1.0, size=(batch, n_points, 3))
    P /= np.max(np.linalg.norm(P, axis=2), axis=1)[:, None, None]
    return P.astype(np.float32)


def lattice_cloud(batch, n_points, seed, pitch=40):
    """Integer lattice / pitch: many exactly equal distances (tie-break stress),
    includes points with |p|^2 <= 1e-3 (the FPS skip rule) and exact duplicates."""
    rng = np.random.default_rng(seed)
    P = rng.integers(-pitch, pitch + 1, size=(batch, n_points, 3)).astype(np.float32) / pitch
    dup = max(1, n_points // 16)
    src = rng.integers(0, n_points, size=(batch, dup))
    dst = rng.integers(0, n_points, size=(batch, dup))
    for b in range(batch):
        P[b, dst[b]] = P[b, src[b]]
    if n_points >= 4:
        P[:, n_points // 3] = 0.0
        P[:, n_points // 2] = np.float32(0.01)
    return P


def network_state(template, seed=7):
    """Deterministic PointNet2 parameters for parity runs: every entry of `template`
    (a reference-layout state dict: name -> tensor/array, only shapes are used) is
    filled from numpy's default_rng, conv weights scaled He-style, BatchNorm with
    non-trivial affine parameters and running statistics so that BN folding is
    exercised.  Returns name -> np.ndarray (num_batches_tracked as int64 zeros)."""
    rng = np.random.default_rng(seed)
    out = {}
    for name in sorted(template.keys()):
        shape = tuple(template[name].shape)
        if name.endswith("num_batches_tracked"):
            out[name] = np.zeros(shape, dtype=np.int64)
        elif name.endswith("running_var"):
            out[name] = rng.uniform(0.5, 1.5, size=shape).astype(np.float32)
        elif name.endswith("running_mean"):
            out[name] = rng.normal(scale=0.1, size=shape).astype(np.float32)
        elif ("bn" in name) and name.endswith("weight"):
            out[name] = rng.uniform(0.8, 1.2, size=shape).astype(np.float32)
        elif name.endswith("bias"):
            out[name] = rng.normal(scale=0.05, size=shape).astype(np.float32)
        else:  # conv weight [Co,Ci,1(,1)]
            fan_in = shape[1]
            out[name] = (rng.normal(size=shape) * np.sqrt(2.0 / fan_in)).astype(np.float32)
    return out


def training_batch(batch, n_points, seed, k_slots=21, n_gt_points=512):
    """Synthetic supervised batch in the layout of Utils/training_utils.py:119-131: P, X_gt [B,N,3], I_gt int64 [B,N]
    (12 ground-truth primitives per cloud), T_gt int64 [B,K] (type ids in the order sphere=0, plane=1, cylinder=2,
    cone=3 of the reference's configuration), points_per_instance [B,K,n_gt_points,3] (points of every ground-truth
    primitive, zeros for the unused slots) and the ground-truth axes plane_normal / cylinder_axis / cone_axis [B,K,3]
    (taken from the analytic normals: exact for planes, a per-primitive mean direction otherwise -- the losses only
    need plausible unit vectors)."""
    P, X, _, I = shape_batch(batch, n_points, seed, k_slots=k_slots)
    rng = np.random.default_rng(seed + 999)
    type_of = np.array([1, 1, 1, 0, 0, 0, 2, 2, 2, 3, 3, 3], dtype=np.int64)      # shape_cloud: 3 planes, 3 spheres, ...
    T_gt = np.zeros((batch, k_slots), np.int64)
    T_gt[:, :12] = type_of
    ppi = np.zeros((batch, k_slots, n_gt_points, 3), np.float32)
    axes = np.zeros((batch, k_slots, 3), np.float32)
    axes[..., 2] = 1.0
    for b in range(batch):
        for g in range(12):
            ids = np.flatnonzero(I[b] == g)
            if len(ids):
                ppi[b, g] = P[b, rng.choice(ids, n_gt_points, replace=len(ids) < n_gt_points)]
                m = X[b, ids].mean(0)
                axes[b, g] = m / max(np.linalg.norm(m), 1e-6)
    return {"P": P.astype(np.float32), "X_gt": X.astype(np.float32), "I_gt": I.astype(np.int64), "T_gt": T_gt,
            "points_per_instance": ppi,
            "gt_parameters": {"plane_normal": axes.copy(), "cylinder_axis": axes.copy(), "cone_axis": axes.copy()}}
