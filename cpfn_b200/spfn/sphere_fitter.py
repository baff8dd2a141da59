"""``compute_parameters(P, W) -> (center [B,K,3], radius_squared [B,K])`` (reference
SPFN/sphere_fitter.py:9-19) and ``compute_residue_single`` (:58-62)."""
import torch

from . import _reference, fit


def sqrt_safe(x):
    return torch.sqrt(torch.abs(x) + 1e-10)


def compute_parameters(P, W):
    r = fit.fit_primitives(P, W, torch.zeros_like(P))
    return r["sphere_center"], r["sphere_radius_squared"]


def compute_residue_single(center, radius_squared, p):
    return (sqrt_safe(torch.sum((p - center) ** 2, dim=-1)) - sqrt_safe(radius_squared)) ** 2


__getattr__ = _reference.forwarder(globals(), "sphere_fitter", ('compute_parameters', 'compute_residue_single', 'sqrt_safe'))
