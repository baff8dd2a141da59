"""numpy fp32 restatement of the reference's point-to-primitive residues.

TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Follows plane_fitter.py:54-55, sphere_fitter.py:61-62,
cylinder_fitter.py:85-89, cone_fitter.py:98-103, losses_implementation.py:351-387 and
metric_implementation.py:76-81, 409-415 line by line.

Parity pin: tests/golden/ref_residues.npz, produced by importing the UNMODIFIED reference SPFN package on CPU
(tests/golden/make_ref_residues_golden.py).
"""
import numpy as np

F = np.float32


def sqrt_safe(x):
    return np.sqrt(np.abs(x) + F(1e-10)).astype(F)


def plane(n, c, p):
    return ((np.sum(p * n, axis=-1, dtype=F) - c) ** 2).astype(F)


def sphere(center, radius_squared, p):
    return ((sqrt_safe(np.sum((p - center) ** 2, axis=-1, dtype=F)) - sqrt_safe(radius_squared)) ** 2).astype(F)


def cylinder(axis, center, radius_squared, p):
    d = p - center
    d2 = np.sum(d ** 2, axis=-1, dtype=F)
    dn = np.sum(d * axis, axis=-1, dtype=F)
    return ((sqrt_safe(d2 - dn ** 2) - sqrt_safe(radius_squared)) ** 2).astype(F)


def cone(apex, axis, half_angle, p):
    v = (p - apex).astype(F)
    nrm = np.maximum(np.sqrt(np.sum(v * v, axis=-1, keepdims=True, dtype=F)), F(1e-12))
    c = np.clip(np.sum((v / nrm) * axis, axis=-1, dtype=F), F(-1.0 + 1e-6), F(1.0 - 1e-6))
    alpha = np.arccos(c).astype(F)
    return (np.sin(np.minimum(np.abs(alpha - half_angle), F(np.pi / 2))) ** 2 * np.sum(v * v, axis=-1, dtype=F)).astype(F)


def _g(t, m):
    return np.take_along_axis(t, m[..., None] if t.ndim == 3 else m, axis=1)


def compute_residue_loss(parameters, matching_indices, points_per_instance, T_gt, classes=('plane', 'sphere', 'cylinder', 'cone')):
    """losses_implementation.py:351-387 -> (residue_loss [B,K], residue_per_point_array [B,K,N',T])."""
    m, p, out = matching_indices, points_per_instance.astype(F), []
    for class_ in classes:
        if class_ == 'plane':
            r = plane(_g(parameters['plane_normal'], m)[:, :, None], _g(parameters['plane_center'], m)[:, :, None], p)
        elif class_ == 'sphere':
            r = sphere(_g(parameters['sphere_center'], m)[:, :, None], _g(parameters['sphere_radius_squared'], m)[:, :, None], p)
        elif class_ == 'cylinder':
            r = cylinder(_g(parameters['cylinder_axis'], m)[:, :, None], _g(parameters['cylinder_center'], m)[:, :, None],
                         _g(parameters['cylinder_radius_squared'], m)[:, :, None], p)
        else:
            r = cone(_g(parameters['cone_apex'], m)[:, :, None], _g(parameters['cone_axis'], m)[:, :, None],
                     _g(parameters['cone_half_angle'], m)[:, :, None], p)
        out.append(r)
    per_point = np.stack(out, axis=3)
    losses = np.stack([np.mean(r, axis=2, dtype=F) for r in out], axis=2)
    return np.take_along_axis(losses, T_gt[..., None], axis=2)[..., 0], per_point


def get_residual_loss(parameters, matching_indices, points_per_instance, T, classes=('plane', 'sphere', 'cylinder', 'cone')):
    """metric_implementation.py:76-81."""
    _, per_point = compute_residue_loss(parameters, matching_indices, points_per_instance,
                                        np.take_along_axis(T, matching_indices, axis=1), classes)
    return sqrt_safe(np.take_along_axis(per_point, T[:, :, None, None], axis=3)[..., 0])


def compute_P_coverage(P, T, matching_indices, parameters, epsilon, classes=('plane', 'sphere', 'cylinder', 'cone')):
    """metric_implementation.py:409-415."""
    B, N, _ = P.shape
    K = T.shape[1]
    r = get_residual_loss(parameters, matching_indices, np.broadcast_to(P[:, None], (B, K, N, 3)),
                          np.take_along_axis(T, matching_indices, axis=1), classes)
    return np.mean((r.min(axis=1) < epsilon).astype(F), axis=1)
