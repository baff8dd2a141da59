"""``PointNet2`` -- the backbone shared by GlobalSPFN, PatchSelection and LocalSPFN, with
the constructor, forward signature, outputs and state-dict layout of the reference
(``PointNet2/pn2_network.py:10-73``): sa1/sa2/sa3, sfp1/sfp2/sfp3, fc1, bn1, fc2.<i>.
A reference checkpoint loads with ``load_state_dict(strict=True)``.

Behaviour kept on purpose: ``F.dropout(p=0.5)`` is called with its default
``training=True`` (reference :63), so the head outputs are stochastic even in
``.eval()``; ``l3_feats`` is the clean global feature.
"""
import torch

from .pointnet2_ops.modules.pointset_abstraction import PointsetAbstraction
from .pointnet2_ops.modules.pointset_feature_propagation import PointsetFeaturePropagation


class PointNet2(torch.nn.Module):
    def __init__(self, dim_input=3, dim_pos=3, output_sizes=[16], use_glob_features=False,
                 use_loc_features=False, features_extractor=False):
        super().__init__()
        self.dim_pos = dim_pos
        self.use_glob_features = use_glob_features
        self.use_loc_features = use_loc_features
        self.features_extractor = features_extractor
        self.sa1 = PointsetAbstraction(num_points=512, dim_pos=dim_pos, dim_feats=dim_input - dim_pos,
                                       radius_list=[0.2], num_samples_list=[64],
                                       mlp_list=[[64, 64, 128]], group_all=False)
        self.sa2 = PointsetAbstraction(num_points=128, dim_pos=dim_pos, dim_feats=128,
                                       radius_list=[0.4], num_samples_list=[64],
                                       mlp_list=[[128, 128, 256]], group_all=False)
        self.sa3 = PointsetAbstraction(num_points=None, dim_pos=dim_pos, dim_feats=256,
                                       radius_list=None, num_samples_list=None,
                                       mlp_list=[256, 512, 1024], group_all=True)
        offset = (1024 if use_glob_features else 0) + (128 if use_loc_features else 0)
        self.sfp1 = PointsetFeaturePropagation(dim_feats=1024 + offset + 256, mlp=[256, 256])
        self.sfp2 = PointsetFeaturePropagation(dim_feats=256 + 128, mlp=[256, 128])
        self.sfp3 = PointsetFeaturePropagation(dim_feats=128 + dim_input - dim_pos, mlp=[128, 128, 128])
        self.fc1 = torch.nn.Conv1d(128, 128, 1)
        if not features_extractor:
            self.bn1 = torch.nn.BatchNorm1d(128)
            self.fc2 = torch.nn.ModuleList(torch.nn.Conv1d(128, n, 1) for n in output_sizes)

    def forward(self, x, glob_features=None, loc_features=None, fast=True):
        """x [B,N,dim_input] -> [head_0 [B,N,o0], ..., l3_feats [B,1024(+),1], output_feat [B,128,N]]
        (or (l3_feats, output_feat) for a features extractor)."""
        x = x.transpose(2, 1)
        input_pos = x[:, :self.dim_pos, :]
        input_feats = x[:, self.dim_pos:, :] if x.shape[1] > self.dim_pos else None
        l1_pos, l1_feats = self.sa1(input_pos, input_feats, fast=fast)
        l2_pos, l2_feats = self.sa2(l1_pos, l1_feats, fast=fast)
        l3_pos, l3_feats = self.sa3(l2_pos, l2_feats, fast=fast)
        if self.use_glob_features:
            l3_feats = torch.cat((l3_feats, glob_features.unsqueeze(2)), dim=1)
        if self.use_loc_features:
            l3_feats = torch.cat((l3_feats, loc_features.unsqueeze(2)), dim=1)
        l4_feats = self.sfp1(l2_pos, l3_pos, l2_feats, l3_feats, fast=fast)
        l5_feats = self.sfp2(l1_pos, l2_pos, l1_feats, l4_feats, fast=fast)
        l6_feats = self.sfp3(input_pos, l1_pos, input_feats, l5_feats, fast=fast)
        output_feat = self.fc1(l6_feats)
        if self.features_extractor:
            return l3_feats, output_feat
        output_feat = torch.nn.functional.relu(self.bn1(output_feat))
        output_feat = torch.nn.functional.dropout(output_feat, p=0.5)
        results = [fc2_layer(output_feat).transpose(1, 2) for fc2_layer in self.fc2]
        results.append(l3_feats)
        results.append(output_feat)
        return results
