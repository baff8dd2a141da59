"""numpy front-end of ``libcpfn_oracle.so`` (see ``cpfn_oracle.c``).

TEST INFRASTRUCTURE.  Shapes and dtypes follow the reference pybind module
``PointNet2/pointnet2_ops/cuda_ops/src/bindings.cpp:6-19``: float32 inputs,
int32 indices, freshly allocated outputs.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libcpfn_oracle.so")
_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int32)


def build(force=False):
    """Compile cpfn_oracle.c with the committed Makefile (gcc, seconds)."""
    src = os.path.join(_HERE, "cpfn_oracle.c")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(src)):
        return _LIB_PATH
    subprocess.run(["make", "-C", _HERE, "-B", "libcpfn_oracle.so"], check=True,
                   stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.cpfn_oracle_opt_n_threads.restype = ctypes.c_int
        _lib.cpfn_oracle_max_threads.restype = ctypes.c_int
    return _lib


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(_f32p)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(_i32p)


def set_threads(t):
    lib().cpfn_oracle_set_threads(int(t))


def max_threads():
    return int(lib().cpfn_oracle_max_threads())


def opt_n_threads(n):
    return int(lib().cpfn_oracle_opt_n_threads(int(n)))


def farthest_point_sampling(points, nsamples):
    """points f32 [B, N, 3] -> i32 [B, nsamples]."""
    points, pp = _f(points)
    B, N, _ = points.shape
    out = np.zeros((B, nsamples), dtype=np.int32)
    lib().cpfn_oracle_fps(B, N, int(nsamples), pp, out.ctypes.data_as(_i32p))
    return out


def ball_query(new_xyz, xyz, radius, nsample):
    """new_xyz f32 [B, S, 3], xyz f32 [B, N, 3] -> i32 [B, S, nsample]."""
    new_xyz, qp = _f(new_xyz)
    xyz, xp = _f(xyz)
    B, S, _ = new_xyz.shape
    N = xyz.shape[1]
    out = np.zeros((B, S, nsample), dtype=np.int32)
    lib().cpfn_oracle_ball_query(B, N, S, ctypes.c_float(radius), int(nsample),
                                 qp, xp, out.ctypes.data_as(_i32p))
    return out


def three_nn(unknown, known):
    """unknown f32 [B, n, 3], known f32 [B, m, 3] -> (dist2 f32, idx i32) [B, n, 3]."""
    unknown, up = _f(unknown)
    known, kp = _f(known)
    B, n, _ = unknown.shape
    m = known.shape[1]
    d2 = np.zeros((B, n, 3), dtype=np.float32)
    idx = np.zeros((B, n, 3), dtype=np.int32)
    lib().cpfn_oracle_three_nn(B, n, m, up, kp, d2.ctypes.data_as(_f32p),
                               idx.ctypes.data_as(_i32p))
    return d2, idx


def three_weighted_sum(points, idx, weight):
    """points f32 [B, C, M], idx i32 [B, n, 3], weight f32 [B, n, 3] -> f32 [B, C, n]."""
    points, pp = _f(points)
    idx, ip = _i(idx)
    weight, wp = _f(weight)
    B, C, M = points.shape
    n = idx.shape[1]
    out = np.zeros((B, C, n), dtype=np.float32)
    lib().cpfn_oracle_three_weighted_sum(B, C, M, n, pp, ip, wp,
                                         out.ctypes.data_as(_f32p))
    return out


def three_weighted_sum_grad(grad_out, idx, weight, m):
    grad_out, gp = _f(grad_out)
    idx, ip = _i(idx)
    weight, wp = _f(weight)
    B, C, n = grad_out.shape
    out = np.zeros((B, C, m), dtype=np.float32)
    lib().cpfn_oracle_three_weighted_sum_grad(B, C, n, int(m), gp, ip, wp,
                                              out.ctypes.data_as(_f32p))
    return out


def gather_points(points, idx):
    points, pp = _f(points)
    idx, ip = _i(idx)
    B, C, N = points.shape
    M = idx.shape[1]
    out = np.zeros((B, C, M), dtype=np.float32)
    lib().cpfn_oracle_gather_points(B, C, N, M, pp, ip, out.ctypes.data_as(_f32p))
    return out


def gather_points_grad(grad_out, idx, n):
    grad_out, gp = _f(grad_out)
    idx, ip = _i(idx)
    B, C, M = grad_out.shape
    out = np.zeros((B, C, n), dtype=np.float32)
    lib().cpfn_oracle_gather_points_grad(B, C, int(n), M, gp, ip,
                                         out.ctypes.data_as(_f32p))
    return out


def group_points(points, idx):
    points, pp = _f(points)
    idx, ip = _i(idx)
    B, C, N = points.shape
    _, S, K = idx.shape
    out = np.zeros((B, C, S, K), dtype=np.float32)
    lib().cpfn_oracle_group_points(B, C, N, S, K, pp, ip, out.ctypes.data_as(_f32p))
    return out


def group_points_grad(grad_out, idx, n):
    grad_out, gp = _f(grad_out)
    idx, ip = _i(idx)
    B, C, S, K = grad_out.shape
    out = np.zeros((B, C, n), dtype=np.float32)
    lib().cpfn_oracle_group_points_grad(B, C, int(n), S, K, gp, ip,
                                        out.ctypes.data_as(_f32p))
    return out
