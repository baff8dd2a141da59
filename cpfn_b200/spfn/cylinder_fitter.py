"""``compute_parameters(P, W, X) -> (axis [B,K,3], center [B,K,3], radius_squared [B,K])``
(reference SPFN/cylinder_fitter.py:10-28) and ``compute_residue_single`` (:82-89)."""
import torch

from . import _reference, fit
from .sphere_fitter import sqrt_safe


def compute_parameters(P, W, X):
    r = fit.fit_primitives(P, W, X)
    return r["cylinder_axis"], r["cylinder_center"], r["cylinder_radius_squared"]


def compute_residue_single(axis, center, radius_squared, p):
    p_minus_c = p - center
    p_minus_c_sqr = torch.sum(p_minus_c ** 2, dim=-1)
    p_minus_c_dot_n = torch.sum(p_minus_c * axis, dim=-1)
    return (sqrt_safe(p_minus_c_sqr - p_minus_c_dot_n ** 2) - sqrt_safe(radius_squared)) ** 2


__getattr__ = _reference.forwarder(globals(), "cylinder_fitter", ('compute_parameters', 'compute_residue_single'))
