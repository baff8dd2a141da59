"""CPU restatement of the GlobalSPFN / LocalSPFN PointNet++ forward (eval mode).

TEST INFRASTRUCTURE, NOT PRODUCT CODE (only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import it).

Follows PointNet2/pn2_network.py:38-73, pointset_abstraction.py:34-76 and
pointset_feature_propagation.py:21-51 of the reference with:
  * the index ops of oracle/cpfn_oracle.c (bit-exact restatement of the reference CUDA
    kernels, pinned to tests/golden/ref_cuda_ops.npz) in place of ``cuda_ops``;
  * torch CPU fp32 ``conv`` / ``batch_norm`` (eval) / ``relu`` / ``max`` for the MLPs,
    exactly the calls the reference modules make;
  * dropout (pn2_network.py:63, always on) replaced by an explicit multiplicative mask
    argument so that a GPU run can be compared with the SAME mask.
Parity pin: the reference network run on CPU tensors needs its CUDA extension for every
``fast=True`` op, so the pin of this file is (a) the pinned index ops and (b) the
reference's own module code for the MLP part, exercised by
tests/golden/make_ref_network_golden.py (fast=False is a different algorithm and is not
used as a pin).
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import index_ops


def _bn(x, sd, prefix):
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"],
                        sd[prefix + ".weight"], sd[prefix + ".bias"], False, 0.1, 1e-5)


def _gather(points, idx):
    """points [B,C,N] torch, idx [B,...] int -> [B,C,...] (select_point_subset)."""
    B, C, _ = points.shape
    flat = torch.as_tensor(np.asarray(idx).reshape(B, -1), dtype=torch.long)
    out = torch.gather(points, 2, flat[:, None, :].expand(B, C, flat.shape[1]))
    return out.view(B, C, *np.asarray(idx).shape[1:])


def set_abstraction(sd, name, pos, feats, num_points, radius, nsample, n_layers, group_all):
    """pos [B,3,N], feats [B,D,N] | None -> (new_pos, new_feats, aux)."""
    B, C, N = pos.shape
    aux = {}
    if group_all:
        grouped = pos.view(B, C, 1, N)
        if feats is not None:
            grouped = torch.cat([grouped, feats.view(B, -1, 1, N)], dim=1)
        new_pos = None
    else:
        xyz = np.ascontiguousarray(pos.permute(0, 2, 1).numpy())
        fps_idx = index_ops.farthest_point_sampling(xyz, num_points)
        new_pos = _gather(pos, fps_idx)
        new_xyz = np.ascontiguousarray(new_pos.permute(0, 2, 1).numpy())
        group_idx = index_ops.ball_query(new_xyz, xyz, radius, nsample)
        aux["fps_idx"], aux["group_idx"] = fps_idx, group_idx
        grouped = _gather(pos, group_idx) - new_pos.view(B, C, num_points, 1)
        if feats is not None:
            grouped = torch.cat([_gather(feats, group_idx), grouped], dim=1)
    for j in range(n_layers):
        w, b = sd["%s.conv_blocks.0.%d.weight" % (name, j)], sd["%s.conv_blocks.0.%d.bias" % (name, j)]
        grouped = F.relu(_bn(F.conv2d(grouped.contiguous(), w, b), sd, "%s.bn_blocks.0.%d" % (name, j)))
    return new_pos, torch.max(grouped, dim=3)[0], aux


def feature_propagation(sd, name, pos1, pos2, feats1, feats2, n_layers):
    B, _, N = pos1.shape
    aux = {}
    if pos2 is None:
        interp = feats2.repeat(1, 1, N)
    else:
        u = np.ascontiguousarray(pos1.permute(0, 2, 1).numpy())
        k = np.ascontiguousarray(pos2.permute(0, 2, 1).numpy())
        d2, idx = index_ops.three_nn(u, k)
        d = torch.sqrt(torch.from_numpy(d2))
        recip = 1.0 / (d + 1e-8)
        w = recip / torch.sum(recip, dim=2, keepdim=True)
        aux["nn_idx"], aux["nn_w"] = idx, w.numpy()
        interp = torch.from_numpy(index_ops.three_weighted_sum(feats2.contiguous().numpy(), idx, w.numpy()))
    x = interp if feats1 is None else torch.cat([feats1, interp], dim=1)
    for j in range(n_layers):
        w, b = sd["%s.mlp_convs.%d.weight" % (name, j)], sd["%s.mlp_convs.%d.bias" % (name, j)]
        x = F.relu(_bn(F.conv1d(x, w, b), sd, "%s.mlp_bns.%d" % (name, j)))
    return x, aux


@torch.no_grad()
def pointnet2_forward(state_dict, x, n_heads, dropout_mask=None):
    """state_dict: reference-layout PointNet2 state dict (CPU fp32 tensors); x [B,N,3+D].
    dropout_mask: None (no dropout, i.e. the expectation) or a [B,128,N] multiplicative mask
    (values 0 or 2).  Returns a dict with every stage output."""
    sd = {k: v.detach().float().cpu() for k, v in state_dict.items()}
    x = torch.as_tensor(np.asarray(x), dtype=torch.float32).transpose(2, 1)
    pos = x[:, :3, :].contiguous()
    feats = x[:, 3:, :].contiguous() if x.shape[1] > 3 else None
    out = {}
    l1_pos, l1_feats, a1 = set_abstraction(sd, "sa1", pos, feats, 512, 0.2, 64, 3, False)
    l2_pos, l2_feats, a2 = set_abstraction(sd, "sa2", l1_pos, l1_feats, 128, 0.4, 64, 3, False)
    _, l3_feats, _ = set_abstraction(sd, "sa3", l2_pos, l2_feats, None, None, None, 3, True)
    l4, _ = feature_propagation(sd, "sfp1", l2_pos, None, l2_feats, l3_feats, 2)
    l5, _ = feature_propagation(sd, "sfp2", l1_pos, l2_pos, l1_feats, l4, 2)
    l6, a6 = feature_propagation(sd, "sfp3", pos, l1_pos, feats, l5, 3)
    feat = F.relu(_bn(F.conv1d(l6, sd["fc1.weight"], sd["fc1.bias"]), sd, "bn1"))
    out.update(sa1_fps=a1["fps_idx"], sa1_group=a1["group_idx"], sa2_fps=a2["fps_idx"],
               sa2_group=a2["group_idx"], fp3_nn=a6["nn_idx"], l1_feats=l1_feats.numpy(),
               l2_feats=l2_feats.numpy(), l3_feats=l3_feats.numpy(), l4_feats=l4.numpy(),
               l5_feats=l5.numpy(), l6_feats=l6.numpy(), feat_pre_dropout=feat.numpy())
    if dropout_mask is not None:
        feat = feat * torch.as_tensor(np.asarray(dropout_mask), dtype=torch.float32)
    out["output_feat"] = feat.numpy()
    out["heads"] = [F.conv1d(feat, sd["fc2.%d.weight" % i], sd["fc2.%d.bias" % i]).transpose(1, 2).numpy()
                    for i in range(n_heads)]
    return out


def spfn_postprocess(heads):
    """Utils/training_utils.py:141-142 of the reference: X normalised, W soft-maxed."""
    X = torch.from_numpy(heads[0])
    X = F.normalize(X, p=2, dim=2, eps=1e-12)
    W = torch.softmax(torch.from_numpy(heads[2]), dim=2)
    return X.numpy(), heads[1], W.numpy()
