"""``PointNet2`` -- the backbone shared by GlobalSPFN, PatchSelection and LocalSPFN.

Interface contract with the reference (``PointNet2/pn2_network.py:10-73``): same constructor
arguments, same ``forward(x, glob_features, loc_features, fast)`` outputs, and the same
sub-module names -- hence the same state-dict keys, so a reference checkpoint loads with
``load_state_dict(strict=True)`` (125 entries for three heads).  The architecture itself is a
table here (``ENCODER`` / ``DECODER``); the layers are the fused-kernel modules of this package.

Kept on purpose: the reference calls ``F.dropout(p=0.5)`` with its default ``training=True``
(reference :63), so the head outputs are stochastic even in ``.eval()``; ``l3_feats`` is clean.
"""
import torch
import torch.nn.functional as F
from torch import nn

from .pointnet2_ops.modules.pointset_abstraction import PointsetAbstraction
from .pointnet2_ops.modules.pointset_feature_propagation import PointsetFeaturePropagation

# name, centroids, radius, samples per ball, MLP widths, input feature channels (None: from dim_input)
ENCODER = (
    ("sa1", 512, 0.2, 64, (64, 64, 128), None),
    ("sa2", 128, 0.4, 64, (128, 128, 256), 128),
    ("sa3", None, None, None, (256, 512, 1024), 256),        # group_all
)
GLOBAL_WIDTH, HIDDEN = 1024, 128


class PointNet2(nn.Module):
    def __init__(self, dim_input=3, dim_pos=3, output_sizes=[16], use_glob_features=False,
                 use_loc_features=False, features_extractor=False):
        super().__init__()
        self.dim_pos = dim_pos
        self.use_glob_features = use_glob_features
        self.use_loc_features = use_loc_features
        self.features_extractor = features_extractor
        extra_in = dim_input - dim_pos
        for name, n_centroids, radius, n_samples, widths, feat_in in ENCODER:
            whole_cloud = n_centroids is None
            setattr(self, name, PointsetAbstraction(
                num_points=n_centroids, dim_pos=dim_pos, dim_feats=extra_in if feat_in is None else feat_in,
                radius_list=None if whole_cloud else [radius],
                num_samples_list=None if whole_cloud else [n_samples],
                mlp_list=list(widths) if whole_cloud else [list(widths)], group_all=whole_cloud))
        injected = (GLOBAL_WIDTH if use_glob_features else 0) + (HIDDEN if use_loc_features else 0)
        decoder = (("sfp1", GLOBAL_WIDTH + injected + 256, (256, 256)),
                   ("sfp2", 256 + 128, (256, 128)),
                   ("sfp3", 128 + extra_in, (128, 128, 128)))
        for name, width_in, widths in decoder:
            setattr(self, name, PointsetFeaturePropagation(dim_feats=width_in, mlp=list(widths)))
        self.fc1 = nn.Conv1d(HIDDEN, HIDDEN, 1)
        if not features_extractor:
            self.bn1 = nn.BatchNorm1d(HIDDEN)
            self.fc2 = nn.ModuleList([nn.Conv1d(HIDDEN, width, 1) for width in output_sizes])

    def _whole_network_fused(self, x):
        """Inference on bare positions without injected features: the whole forward is the fused pipeline of
        cpfn_b200/fused.py (side-stream overlap, FP3 + fc1 + dropout + heads as one kernel) instead of one fused
        call per module."""
        from . import fused
        if self.training or self.use_glob_features or self.use_loc_features or self.features_extractor:
            return False
        if self.dim_pos != 3 or x.dim() != 3 or x.shape[2] != 3 or not x.is_cuda or x.dtype != torch.float32:
            return False
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            return False
        return fused.available()

    def forward(self, x, glob_features=None, loc_features=None, fast=True):
        """x [B,N,dim_input] -> [head_0 [B,N,o0], ..., l3_feats [B,1024(+),1], output_feat [B,128,N]]
        (a features extractor returns ``(l3_feats, output_feat)`` with output_feat = fc1 output)."""
        if fast and self._whole_network_fused(x):
            from . import fused
            heads, code, hidden = fused.pointnet2_forward(self, x, dropout=True)[:3]
            return list(heads) + [code, hidden]
        channels_first = x.transpose(2, 1)
        xyz0 = channels_first[:, :self.dim_pos, :]
        f0 = channels_first[:, self.dim_pos:, :] if channels_first.shape[1] > self.dim_pos else None
        xyz1, f1 = self.sa1(xyz0, f0, fast=fast)
        xyz2, f2 = self.sa2(xyz1, f1, fast=fast)
        _, code = self.sa3(xyz2, f2, fast=fast)
        for enabled, injected in ((self.use_glob_features, glob_features), (self.use_loc_features, loc_features)):
            if enabled:
                code = torch.cat((code, injected.unsqueeze(2)), dim=1)
        up2 = self.sfp1(xyz2, None, f2, code, fast=fast)
        up1 = self.sfp2(xyz1, xyz2, f1, up2, fast=fast)
        up0 = self.sfp3(xyz0, xyz1, f0, up1, fast=fast)
        hidden = self.fc1(up0)
        if self.features_extractor:
            return code, hidden
        hidden = F.dropout(F.relu(self.bn1(hidden)), p=0.5)          # always on, as in the reference
        return [head(hidden).transpose(1, 2) for head in self.fc2] + [code, hidden]
