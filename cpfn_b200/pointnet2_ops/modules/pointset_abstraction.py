"""``PointsetAbstraction``: sample centroids, group their neighbourhoods, run a shared MLP on every
neighbourhood and max-pool it -- the set-abstraction layer of PointNet++.

Interface contract with the reference (``PointNet2/pointnet2_ops/modules/pointset_abstraction.py:7-76``):
constructor arguments, ``forward(pos, feats, fast)`` and the parameter names
``conv_blocks.<scale>.<layer>.*`` / ``bn_blocks.<scale>.<layer>.*`` (reference checkpoints load with
``strict=True``).  Channel order of a grouped row: ``[feats..., dx, dy, dz]`` for ball grouping,
``[x, y, z, feats...]`` without re-centring for ``group_all`` (reference :54-56, :62-66).

Two execution paths, both on libcpfn_b200.so kernels:
  * inference (``eval()`` and no autograd): ONE fused kernel per scale gathers the neighbourhoods into
    shared memory, runs the BatchNorm-folded MLP on the tcgen05 tensor cores and max-pools in
    registers (``cpfn_b200/fused.py``): the grouped [B,C,S,K] tensor never exists in HBM;
  * training / autograd: per-op kernels (gather / group with scatter-add backward) and torch's conv /
    batch-norm modules, because training-mode BatchNorm needs batch statistics between the layers.
"""
from collections.abc import Sequence

import torch
import torch.nn.functional as F
from torch import nn

from . import geometry_utils as G
from ... import fused


def _as_list(v):
    return list(v) if isinstance(v, Sequence) else [v]


class PointsetAbstraction(nn.Module):
    def __init__(self, num_points, dim_pos, dim_feats, radius_list, num_samples_list, mlp_list, group_all=False):
        super().__init__()
        self.num_points, self.group_all = num_points, group_all
        self.radius_list, self.num_samples_list = _as_list(radius_list), _as_list(num_samples_list)
        self.mlp_list = mlp_list if isinstance(mlp_list[0], Sequence) else [mlp_list]
        if not (len(self.radius_list) == len(self.num_samples_list) == len(self.mlp_list)):
            raise ValueError('Radius, number of samples and mlps lists must have the same number of entries.')
        self.conv_blocks, self.bn_blocks = nn.ModuleList(), nn.ModuleList()
        for widths in self.mlp_list:                                  # one shared MLP per grouping scale
            dims = [dim_pos + dim_feats] + list(widths)
            self.conv_blocks.append(nn.ModuleList(nn.Conv2d(a, b, 1) for a, b in zip(dims[:-1], dims[1:])))
            self.bn_blocks.append(nn.ModuleList(nn.BatchNorm2d(b) for b in dims[1:]))

    def _inference(self, pos, feats):
        if self.training or pos.shape[1] != 3 or not fused.available():
            return False
        tracked = [pos, feats] + list(self.parameters())
        return not (torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tracked))

    def _grouped_rows(self, scale, pos, feats, xyz, new_pos, new_xyz):
        """[B, C(+D), S, K] rows of one scale (training path)."""
        B, C, N = pos.shape
        if self.group_all:
            rows = pos.reshape(B, C, 1, N)
            return rows if feats is None else torch.cat([rows, feats.reshape(B, -1, 1, N)], dim=1)
        idx = G.ball_query_nc(self.radius_list[scale], self.num_samples_list[scale], xyz, new_xyz)   # int32 [B,S,K]
        rows = G.select_point_subset(pos, idx) - new_pos.unsqueeze(-1)
        return rows if feats is None else torch.cat([G.select_point_subset(feats, idx), rows], dim=1)

    def forward(self, pos, feats, fast=True):
        """pos [B,C,N], feats [B,D,N] | None -> (new_pos [B,C,S] | None, new_feats [B,D',S])."""
        if not fast:
            G._no_slow_path("PointsetAbstraction")
        if self._inference(pos, feats):
            return fused.set_abstraction_forward(self, pos, feats)
        xyz = new_pos = new_xyz = None
        if not self.group_all:
            xyz = pos.detach().permute(0, 2, 1).contiguous()                       # [B,N,3]
            new_pos = G.select_point_subset(pos, G.farthest_point_sample_nc(xyz, self.num_points))
            new_xyz = new_pos.detach().permute(0, 2, 1).contiguous()
        pooled = []
        for scale, (convs, bns) in enumerate(zip(self.conv_blocks, self.bn_blocks)):
            h = self._grouped_rows(scale, pos, feats, xyz, new_pos, new_xyz)
            for conv, bn in zip(convs, bns):
                h = F.relu(bn(conv(h.contiguous())))
            pooled.append(h.max(dim=3)[0])
        return new_pos, torch.cat(pooled, dim=1)
