"""``PointsetAbstraction`` with the constructor, forward signature and state-dict keys
of the reference (``PointNet2/pointnet2_ops/modules/pointset_abstraction.py:7-76``) so
reference checkpoints load with ``strict=True``:
``conv_blocks.<i>.<j>.{weight[Co,Ci,1,1],bias}``, ``bn_blocks.<i>.<j>.*``.

Forward = furthest point sampling -> centroid gather -> ball query -> grouping with
re-centring -> shared MLP (1x1 conv + BatchNorm + ReLU) x n -> max over the samples.
Two execution paths, both on libcpfn_b200.so kernels:
  * inference (``not self.training`` and no autograd): ONE fused kernel per scale
    gathers the neighbourhoods into shared memory, runs the BN-folded MLP on the
    tcgen05 tensor cores and max-pools in registers (cpfn_b200/fused.py) -- the grouped
    [B,C,S,K] tensor never exists in HBM;
  * training / autograd: per-op kernels (gather / group with scatter-add backward) and
    torch's conv / batch-norm modules, because training-mode BatchNorm needs batch
    statistics between the layers.
"""
from collections.abc import Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import geometry_utils as G
from ... import fused


class PointsetAbstraction(nn.Module):
    def __init__(self, num_points, dim_pos, dim_feats, radius_list, num_samples_list, mlp_list,
                 group_all=False):
        super().__init__()
        self.num_points = num_points
        self.group_all = group_all
        self.radius_list = radius_list if isinstance(radius_list, Sequence) else [radius_list]
        self.num_samples_list = (num_samples_list if isinstance(num_samples_list, Sequence)
                                 else [num_samples_list])
        self.mlp_list = mlp_list if isinstance(mlp_list[0], Sequence) else [mlp_list]
        if (len(self.radius_list) != len(self.num_samples_list)
                or len(self.radius_list) != len(self.mlp_list)):
            raise ValueError('Radius, number of samples and mlps lists must have the same number of entries.')
        self.conv_blocks = nn.ModuleList()
        self.bn_blocks = nn.ModuleList()
        for mlp in self.mlp_list:
            convs, bns = nn.ModuleList(), nn.ModuleList()
            cin = dim_pos + dim_feats
            for cout in mlp:
                convs.append(nn.Conv2d(cin, cout, 1))
                bns.append(nn.BatchNorm2d(cout))
                cin = cout
            self.conv_blocks.append(convs)
            self.bn_blocks.append(bns)

    def _use_fused(self, pos, feats):
        needs_grad = torch.is_grad_enabled() and (
            pos.requires_grad or (feats is not None and feats.requires_grad)
            or any(p.requires_grad for p in self.parameters()))
        return (not self.training) and (not needs_grad) and pos.shape[1] == 3 and fused.available()

    def forward(self, pos, feats, fast=True):
        """pos [B,C,N], feats [B,D,N] | None -> (new_pos [B,C,S] | None, new_feats [B,D',S])."""
        if not fast:
            G._no_slow_path("PointsetAbstraction")
        B, C, N = pos.shape
        S = self.num_points
        if self._use_fused(pos, feats):
            return fused.set_abstraction_forward(self, pos, feats)
        if self.group_all:
            new_pos = None
        else:
            xyz = pos.detach().permute(0, 2, 1).contiguous()            # [B,N,3]
            fps_idx = G.farthest_point_sample_nc(xyz, S)                # int32 [B,S]
            new_pos = G.select_point_subset(pos, fps_idx)               # [B,C,S]
            new_xyz = new_pos.detach().permute(0, 2, 1).contiguous()
        new_feats_list = []
        for i, r in enumerate(self.radius_list):
            if self.group_all:
                grouped = pos.reshape(B, C, 1, N)
                if feats is not None:
                    grouped = torch.cat([grouped, feats.reshape(B, -1, 1, N)], dim=1)
            else:
                K = self.num_samples_list[i]
                group_idx = G.ball_query_nc(r, K, xyz, new_xyz)         # int32 [B,S,K]
                grouped = G.select_point_subset(pos, group_idx) - new_pos.view(B, C, S, 1)
                if feats is not None:
                    grouped = torch.cat([G.select_point_subset(feats, group_idx), grouped], dim=1)
            for conv, bn in zip(self.conv_blocks[i], self.bn_blocks[i]):
                grouped = F.relu(bn(conv(grouped.contiguous())))
            new_feats_list.append(torch.max(grouped, dim=3)[0])
        return new_pos, torch.cat(new_feats_list, dim=1)
