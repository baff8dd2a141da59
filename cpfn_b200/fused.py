"""Fused inference kernels behind PointsetAbstraction / PointsetFeaturePropagation.

Placeholder until the tcgen05 kernels land: ``available()`` is False and the modules
use the per-op kernels.
"""


def available():
    return False


def set_abstraction_forward(module, pos, feats):
    raise RuntimeError("fused set abstraction is not built")


def feature_propagation_forward(module, pos1, pos2, feats1, feats2):
    raise RuntimeError("fused feature propagation is not built")


def invalidate(model):
    """Drop cached folded weights after a state-dict load."""
    return None
