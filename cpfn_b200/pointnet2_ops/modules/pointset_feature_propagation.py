"""``PointsetFeaturePropagation`` with the constructor, forward signature and state-dict
keys (``mlp_convs.<i>.*``, ``mlp_bns.<i>.*``) of the reference
(``PointNet2/pointnet2_ops/modules/pointset_feature_propagation.py:6-51``).

Forward = 3-NN -> inverse-distance weights 1/(d+1e-8) normalised (d is the sqrt
distance, as on the reference's CUDA path) -> weighted sum -> concat with the skip
features -> (1x1 conv + BatchNorm + ReLU) x n.  In inference the interpolation, the
concat and the BN-folded MLP run as one fused tcgen05 kernel (cpfn_b200/fused.py);
with autograd the per-op kernels + torch modules are used (training-mode BatchNorm).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import geometry_utils as G
from ... import fused


class PointsetFeaturePropagation(nn.Module):
    def __init__(self, dim_feats, mlp):
        super().__init__()
        self.mlp_convs = nn.ModuleList()
        self.mlp_bns = nn.ModuleList()
        cin = dim_feats
        for cout in mlp:
            self.mlp_convs.append(nn.Conv1d(cin, cout, 1))
            self.mlp_bns.append(nn.BatchNorm1d(cout))
            cin = cout

    def _use_fused(self, *tensors):
        needs_grad = torch.is_grad_enabled() and (
            any(t is not None and t.requires_grad for t in tensors)
            or any(p.requires_grad for p in self.parameters()))
        return (not self.training) and (not needs_grad) and fused.available()

    def forward(self, pos1, pos2, feats1, feats2, fast=True):
        """pos1 [B,C,N], pos2 [B,C,S] | None, feats1 [B,D,N] | None, feats2 [B,D2,S] -> [B,D',N]."""
        if not fast:
            G._no_slow_path("PointsetFeaturePropagation")
        B, _, N = pos1.shape
        if self._use_fused(pos1, pos2, feats1, feats2):
            return fused.feature_propagation_forward(self, pos1, pos2, feats1, feats2)
        if pos2 is None:
            interpolated = feats2.repeat(1, 1, N)
        else:
            dists, idx = G.three_nn_nc(pos1.detach().permute(0, 2, 1).contiguous(),
                                       pos2.detach().permute(0, 2, 1).contiguous())
            recip = 1.0 / (dists + 1e-8)
            weights = recip / torch.sum(recip, dim=2, keepdim=True)
            interpolated = G.three_weighted_sum(feats2, idx, weights)
        new_feats = interpolated if feats1 is None else torch.cat([feats1, interpolated], dim=1)
        for conv, bn in zip(self.mlp_convs, self.mlp_bns):
            new_feats = F.relu(bn(conv(new_feats)))
        return new_feats
