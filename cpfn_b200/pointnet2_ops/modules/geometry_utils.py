"""Functional point-set API with the names, argument order and return types of the
reference ``PointNet2/pointnet2_ops/modules/geometry_utils.py`` (farthest_point_sample
:74-101, ball_query :133-161, three_nn :192-215, three_weighted_sum :267-283,
select_point_subset :26-44, pairwise_squared_distance :4-24).

Every function runs the sm_100a kernels of libcpfn_b200.so through
``cpfn_b200.cuda_ops``.  The reference's ``fast=False`` branch is a different
algorithm (random FPS seed, ``>`` instead of ``<`` at the ball boundary,
matmul-expanded distances, squared instead of sqrt 3-NN distances; SURVEY.md
section 4) and is not part of the hot path: ``fast=False`` raises.
Layout: positions are [B,3,N] at this API (as in the reference); the ``*_nc``
helpers take the kernels' native [B,N,3] float32 / int32 tensors and skip the
permute / int64 round trips.
"""
import torch

from ... import cuda_ops


def _no_slow_path(name):
    raise NotImplementedError(
        "cpfn_b200.%s: fast=False (the reference's pure-torch composition) is not "
        "implemented; this package is the CUDA hot path only" % name)


def pairwise_squared_distance(src, dst):
    """src [B,C,N], dst [B,C,M] -> [B,N,M] (reference :4-24, same expansion)."""
    B, _, N = src.shape
    M = dst.shape[2]
    dist = -2 * torch.matmul(src.permute(0, 2, 1), dst)
    dist += torch.sum(src ** 2, dim=1).view(B, N, 1)
    dist += torch.sum(dst ** 2, dim=1).view(B, 1, M)
    return dist


class _Gather(torch.autograd.Function):
    """points [B,C,N], idx int32 [B,M] or [B,S,K] -> [B,C,M] / [B,C,S,K]; backward is
    the scatter-add kernel (cpfn_gather_points_grad / cpfn_group_points_grad)."""

    @staticmethod
    def forward(ctx, points, idx):
        ctx.save_for_backward(idx)
        ctx.n = points.size(2)
        if idx.dim() == 2:
            return cuda_ops.gather_points(points, idx)
        return cuda_ops.group_points(points, idx)

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        grad_out = grad_out.contiguous()
        if idx.dim() == 2:
            return cuda_ops.gather_points_grad(grad_out, idx, ctx.n), None
        return cuda_ops.group_points_grad(grad_out, idx, ctx.n), None


def select_point_subset(points, idx):
    """points [B,C,N], idx [B,*] (int64 or int32) -> [B,C,*] (reference :26-44)."""
    if idx.dim() not in (2, 3):
        lead = idx.shape[1:]
        out = select_point_subset(points, idx.reshape(idx.size(0), -1))
        return out.view(points.size(0), points.size(1), *lead)
    pts = points if points.dtype == torch.float32 else points.float()
    out = _Gather.apply(pts.contiguous(), idx.to(torch.int32).contiguous())
    return out if out.dtype == points.dtype else out.to(points.dtype)


# ---- native-layout helpers ([B,N,3] positions, int32 indices) ---------------------------------

def farthest_point_sample_nc(xyz, num_point):
    """xyz f32 [B,N,3] contiguous -> int32 [B,num_point]."""
    return cuda_ops.farthest_point_sampling(xyz, num_point)


def ball_query_nc(radius, num_samples, xyz, new_xyz):
    """xyz [B,N,3], new_xyz [B,S,3] -> int32 [B,S,num_samples]."""
    return cuda_ops.ball_query(new_xyz, xyz, radius, num_samples)


def three_nn_nc(unknown, known):
    """unknown [B,n,3], known [B,m,3] -> (dist f32 [B,n,3] (sqrt), idx int32 [B,n,3])."""
    dist2, idx = cuda_ops.three_nn(unknown, known)
    return torch.sqrt(dist2), idx


# ---- reference-named API ----------------------------------------------------------------------

def farthest_point_sample(point_pos, num_point, fast=True):
    """point_pos [B,3,N] -> int64 [B,num_point] (reference :74-86)."""
    if not fast:
        _no_slow_path("farthest_point_sample")
    if point_pos.shape[1] != 3:
        raise ValueError('Points must have exactly three position dimensions when using the fast method.')
    xyz = point_pos.detach().permute(0, 2, 1).contiguous()
    return farthest_point_sample_nc(xyz, num_point).to(dtype=torch.long)


def ball_query(radius, num_samples, point_pos, query_pos, fast=True):
    """point_pos [B,3,N], query_pos [B,3,S] -> int64 [B,S,num_samples] (reference :133-149)."""
    if not fast:
        _no_slow_path("ball_query")
    if point_pos.shape[1] != 3:
        raise ValueError('Points must have exactly three position dimensions when using the fast method.')
    return ball_query_nc(radius, num_samples, point_pos.detach().permute(0, 2, 1).contiguous(),
                         query_pos.detach().permute(0, 2, 1).contiguous()).to(dtype=torch.long)


def three_nn(point_pos, query_pos, fast=True):
    """point_pos [B,3,N], query_pos [B,3,S] -> (dists [B,S,3] (sqrt, as the reference's
    fast path :184), indices int64 [B,S,3]) (reference :192-210)."""
    if not fast:
        _no_slow_path("three_nn")
    if point_pos.shape[1] != 3:
        raise ValueError('Points must have exactly three position dimensions when using the fast method.')
    dists, indices = three_nn_nc(query_pos.detach().permute(0, 2, 1).contiguous(),
                                 point_pos.detach().permute(0, 2, 1).contiguous())
    return dists, indices.to(dtype=torch.long)


class _ThreeWeightedSum(torch.autograd.Function):
    """reference _FastThreeWeightedSum (:217-263): gradient flows to the features only."""

    @staticmethod
    def forward(ctx, features, idx, weight):
        idx = idx.to(torch.int32).contiguous()
        weight = weight.contiguous()
        ctx.save_for_backward(idx, weight)
        ctx.n = features.size(2)
        return cuda_ops.three_weighted_sum(features.contiguous(), idx, weight)

    @staticmethod
    def backward(ctx, grad_out):
        idx, weight = ctx.saved_tensors
        return cuda_ops.three_weighted_sum_grad(grad_out.contiguous(), idx, weight, ctx.n), None, None


def three_weighted_sum(point_feats, indices, weights, fast=True):
    """point_feats [B,C,N], indices [B,S,3], weights [B,S,3] -> [B,C,S] (reference :267-283)."""
    if not fast:
        _no_slow_path("three_weighted_sum")
    return _ThreeWeightedSum.apply(point_feats, indices, weights)
