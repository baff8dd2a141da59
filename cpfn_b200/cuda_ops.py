"""Drop-in replacement of the reference pybind module ``cuda_ops``.

Same nine functions, positional signatures, dtypes, shapes and error behaviour
as ``PointNet2/pointnet2_ops/cuda_ops/src/bindings.cpp:6-19`` (checks from
``include/utils.h:5-25``): inputs must be contiguous float32 / int32 CUDA
tensors, outputs are freshly allocated on the input's device, work is enqueued
on torch's current CUDA stream, CPU tensors raise ``RuntimeError("CPU not
supported")``.  Each call forwards raw device pointers to the C ABI in
``include/cpfn_b200.h``; torch is only the allocator and the stream provider.
"""
import os

import torch

from . import _lib

def _bq_workspace(nbytes, device):
    """Grid workspace of the uniform-grid ball query: a fresh block from torch's caching allocator per call
    (stream-safe; a captured CUDA graph keeps the block it was captured with alive)."""
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


# Number of kernels this package enqueued (bench.py reports it as gpu_launches).
LAUNCHES = 0


def count_launches(n=1):
    global LAUNCHES
    LAUNCHES += n


def _check_contiguous(x, name):
    if not x.is_contiguous():
        raise RuntimeError("%s must be a contiguous tensor" % name)


def _check_float(x, name):
    if x.dtype != torch.float32:
        raise RuntimeError("%s must be a float tensor" % name)


def _check_int(x, name):
    if x.dtype != torch.int32:
        raise RuntimeError("%s must be an int tensor" % name)


def _check_cuda(x, name):
    if not x.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor" % name)


def _need_cuda(x):
    if not x.is_cuda:
        raise RuntimeError("CPU not supported")


def _stream(x):
    return torch.cuda.current_stream(x.device).cuda_stream


def _p(x):
    return x.data_ptr()


def _check(code, what, launches=1):
    _lib.check(code, what)
    count_launches(launches)


def farthest_point_sampling(points, nsamples, return_centroids=False):
    """points f32 [B,N,3] -> i32 [B,nsamples] (sampling.cpp:65-86).  ``return_centroids`` (an extension, off by
    default) also returns the sampled points f32 [B,nsamples,3], written by the same kernel."""
    _check_contiguous(points, "points")
    _check_float(points, "points")
    _need_cuda(points)
    B, N = points.size(0), points.size(1)
    out = torch.empty((B, nsamples), dtype=torch.int32, device=points.device)
    cen = torch.empty((B, nsamples, 3), dtype=torch.float32, device=points.device) if return_centroids else None
    L = _lib.lib()
    ws_bytes = L.cpfn_fps_workspace_bytes(B, N)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=points.device) if ws_bytes else None
    with torch.cuda.device(points.device):
        _check(L.cpfn_furthest_point_sampling_xyz(_p(points), B, N, int(nsamples), _p(out),
                                                  _p(cen) if cen is not None else None,
                                                  _p(ws) if ws is not None else None, ws_bytes,
                                                  _stream(points)), "farthest_point_sampling")
    return (out, cen) if return_centroids else out


def ball_query(new_xyz, xyz, radius, nsample):
    """new_xyz f32 [B,S,3], xyz f32 [B,N,3] -> i32 [B,S,nsample] (ball_query.cpp:8-32)."""
    _check_contiguous(new_xyz, "new_xyz")
    _check_contiguous(xyz, "xyz")
    _check_float(new_xyz, "new_xyz")
    _check_float(xyz, "xyz")
    if new_xyz.is_cuda:
        _check_cuda(xyz, "xyz")
    _need_cuda(new_xyz)
    B, S = new_xyz.size(0), new_xyz.size(1)
    N = xyz.size(1)
    out = torch.empty((B, S, nsample), dtype=torch.int32, device=new_xyz.device)
    L = _lib.lib()
    with torch.cuda.device(new_xyz.device):
        if 2048 <= N <= 32768 and float(radius) > 0 and not os.environ.get("CPFN_BQ_NO_GRID"):
            # uniform-grid kernel (bit-identical result, ~10x fewer distance tests)
            nbytes = L.cpfn_ball_query_grid_workspace_bytes(xyz.size(0), N)
            ws = _bq_workspace(nbytes, new_xyz.device)
            _check(L.cpfn_ball_query_grid(_p(new_xyz), _p(xyz), xyz.size(0), N, S, float(radius), int(nsample), _p(out),
                                          _p(ws), ws.numel(), _stream(new_xyz)), "ball_query", launches=2)
        else:
            _check(L.cpfn_ball_query(_p(new_xyz), _p(xyz), xyz.size(0), N, S, float(radius), int(nsample), _p(out),
                                     _stream(new_xyz)), "ball_query")
    return out


def gather_points(points, idx):
    """points f32 [B,C,N], idx i32 [B,M] -> f32 [B,C,M] (sampling.cpp:15-40)."""
    _check_contiguous(points, "points")
    _check_contiguous(idx, "idx")
    _check_float(points, "points")
    _check_int(idx, "idx")
    if points.is_cuda:
        _check_cuda(idx, "idx")
    _need_cuda(points)
    B, C, N = points.shape
    M = idx.size(1)
    out = torch.empty((B, C, M), dtype=torch.float32, device=points.device)
    with torch.cuda.device(points.device):
        _check(_lib.lib().cpfn_gather_points(_p(points), _p(idx), B, C, N, M, _p(out),
                                                 _stream(points)), "gather_points")
    return out


def gather_points_grad(grad_out, idx, n):
    """grad_out f32 [B,C,M], idx i32 [B,M] -> f32 [B,C,n] (sampling.cpp:42-64)."""
    _check_contiguous(grad_out, "grad_out")
    _check_contiguous(idx, "idx")
    _check_float(grad_out, "grad_out")
    _check_int(idx, "idx")
    if grad_out.is_cuda:
        _check_cuda(idx, "idx")
    _need_cuda(grad_out)
    B, C, M = grad_out.shape
    out = torch.empty((B, C, n), dtype=torch.float32, device=grad_out.device)
    with torch.cuda.device(grad_out.device):
        _check(_lib.lib().cpfn_gather_points_grad(_p(grad_out), _p(idx), B, C, int(n), M,
                                                      _p(out), _stream(grad_out)),
                   "gather_points_grad")
    return out


def group_points(points, idx):
    """points f32 [B,C,N], idx i32 [B,S,K] -> f32 [B,C,S,K] (group_points.cpp:12-35)."""
    _check_contiguous(points, "points")
    _check_contiguous(idx, "idx")
    _check_float(points, "points")
    _check_int(idx, "idx")
    if points.is_cuda:
        _check_cuda(idx, "idx")
    _need_cuda(points)
    B, C, N = points.shape
    S, K = idx.size(1), idx.size(2)
    out = torch.empty((B, C, S, K), dtype=torch.float32, device=points.device)
    with torch.cuda.device(points.device):
        _check(_lib.lib().cpfn_group_points(_p(points), _p(idx), B, C, N, S, K, _p(out),
                                                _stream(points)), "group_points")
    return out


def group_points_grad(grad_out, idx, n):
    """grad_out f32 [B,C,S,K], idx i32 [B,S,K] -> f32 [B,C,n] (group_points.cpp:37-60)."""
    _check_contiguous(grad_out, "grad_out")
    _check_contiguous(idx, "idx")
    _check_float(grad_out, "grad_out")
    _check_int(idx, "idx")
    if grad_out.is_cuda:
        _check_cuda(idx, "idx")
    _need_cuda(grad_out)
    B, C = grad_out.size(0), grad_out.size(1)
    S, K = idx.size(1), idx.size(2)
    out = torch.empty((B, C, n), dtype=torch.float32, device=grad_out.device)
    with torch.cuda.device(grad_out.device):
        _check(_lib.lib().cpfn_group_points_grad(_p(grad_out), _p(idx), B, C, int(n), S, K,
                                                     _p(out), _stream(grad_out)),
                   "group_points_grad")
    return out


def three_nn(unknowns, knows):
    """unknowns f32 [B,n,3], knows f32 [B,m,3] -> [dist2 f32 [B,n,3], idx i32 [B,n,3]]
    (interpolate.cpp:14-40; distances are SQUARED, the caller takes the sqrt)."""
    _check_contiguous(unknowns, "unknowns")
    _check_contiguous(knows, "knows")
    _check_float(unknowns, "unknowns")
    _check_float(knows, "knows")
    if unknowns.is_cuda:
        _check_cuda(knows, "knows")
    _need_cuda(unknowns)
    B, n = unknowns.size(0), unknowns.size(1)
    m = knows.size(1)
    idx = torch.empty((B, n, 3), dtype=torch.int32, device=unknowns.device)
    dist2 = torch.empty((B, n, 3), dtype=torch.float32, device=unknowns.device)
    with torch.cuda.device(unknowns.device):
        _check(_lib.lib().cpfn_three_nn(_p(unknowns), _p(knows), B, n, m, _p(dist2), _p(idx),
                                            _stream(unknowns)), "three_nn")
    return [dist2, idx]


def three_weighted_sum(points, idx, weight):
    """points f32 [B,C,M], idx i32 [B,n,3], weight f32 [B,n,3] -> f32 [B,C,n]
    (interpolate.cpp:42-70)."""
    _check_contiguous(points, "points")
    _check_contiguous(idx, "idx")
    _check_contiguous(weight, "weight")
    _check_float(points, "points")
    _check_int(idx, "idx")
    _check_float(weight, "weight")
    if points.is_cuda:
        _check_cuda(idx, "idx")
        _check_cuda(weight, "weight")
    _need_cuda(points)
    B, C, M = points.shape
    n = idx.size(1)
    out = torch.empty((B, C, n), dtype=torch.float32, device=points.device)
    with torch.cuda.device(points.device):
        _check(_lib.lib().cpfn_three_weighted_sum(_p(points), _p(idx), _p(weight), B, C, M, n,
                                                      _p(out), _stream(points)),
                   "three_weighted_sum")
    return out


def three_weighted_sum_grad(grad_out, idx, weight, m):
    """grad_out f32 [B,C,n], idx i32 [B,n,3], weight f32 [B,n,3] -> f32 [B,C,m]
    (interpolate.cpp:71-99)."""
    _check_contiguous(grad_out, "grad_out")
    _check_contiguous(idx, "idx")
    _check_contiguous(weight, "weight")
    _check_float(grad_out, "grad_out")
    _check_int(idx, "idx")
    _check_float(weight, "weight")
    if grad_out.is_cuda:
        _check_cuda(idx, "idx")
        _check_cuda(weight, "weight")
    _need_cuda(grad_out)
    B, C, n = grad_out.shape
    out = torch.empty((B, C, m), dtype=torch.float32, device=grad_out.device)
    with torch.cuda.device(grad_out.device):
        _check(_lib.lib().cpfn_three_weighted_sum_grad(_p(grad_out), _p(idx), _p(weight), B, C,
                                                           n, int(m), _p(out), _stream(grad_out)),
                   "three_weighted_sum_grad")
    return out
