"""``compute_parameters(P, W) -> (n [B,K,3], c [B,K])`` (reference SPFN/plane_fitter.py:9-17)
and ``compute_residue_single`` (:54-55)."""
import torch

from . import _reference, fit


def compute_parameters(P, W):
    r = fit.fit_primitives(P, W, torch.zeros_like(P))
    return r["plane_normal"], r["plane_center"]


def compute_residue_single(n, c, p):
    return (torch.sum(p * n, dim=-1) - c) ** 2


__getattr__ = _reference.forwarder(globals(), "plane_fitter", ('compute_parameters', 'compute_residue_single'))
