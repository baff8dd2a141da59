"""``compute_parameters(P, W, X) -> (apex [B,K,3], axis [B,K,3], half_angle [B,K])``
(reference SPFN/cone_fitter.py:12-36) and ``compute_residue_single`` (:98-103)."""
import math

import torch

from . import _reference, fit


def acos_safe(x):
    return torch.acos(torch.clamp(x, min=-1.0 + 1e-6, max=1.0 - 1e-6))


def compute_parameters(P, W, X, div_eps=1e-10):
    r = fit.fit_primitives(P, W, X)
    return r["cone_apex"], r["cone_axis"], r["cone_half_angle"]


def compute_residue_single(apex, axis, half_angle, p):
    v = p - apex
    v_normalized = torch.nn.functional.normalize(v, p=2, dim=-1, eps=1e-12)
    alpha = acos_safe(torch.sum(v_normalized * axis, dim=-1))
    return (torch.sin(torch.clamp(torch.abs(alpha - half_angle), max=math.pi / 2))) ** 2 * torch.sum(v * v, dim=-1)


__getattr__ = _reference.forwarder(globals(), "cone_fitter", ('compute_parameters', 'compute_residue_single', 'acos_safe'))
