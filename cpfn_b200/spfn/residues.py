"""Fused point-to-primitive residues on the GPU (csrc/residues.cu).

``residues(parameters, matching_indices, points, classes)`` evaluates, in one kernel, what the reference
computes with four ``compute_residue_single`` chains of element-wise torch kernels inside
``compute_residue_loss`` (SPFN/losses_implementation.py:351-387); ``p_coverage`` is
``compute_P_coverage`` (SPFN/metric_implementation.py:409-415) without its [B,K,N,T] intermediate.
Inference / metric path only (no autograd): ``losses_implementation.compute_residue_loss`` switches to the
element-wise torch restatement when a gradient is required."""
import ctypes

import torch

from .. import _lib, cuda_ops

CLASS_ID = {"plane": 0, "sphere": 1, "cylinder": 2, "cone": 3}
_FIELDS = ("plane_normal", "plane_center", "sphere_center", "sphere_radius_squared", "cylinder_axis",
           "cylinder_center", "cylinder_radius_squared", "cone_apex", "cone_axis", "cone_half_angle")
_NEEDS = {0: _FIELDS[0:2], 1: _FIELDS[2:4], 2: _FIELDS[4:7], 3: _FIELDS[7:10]}


class _Params(ctypes.Structure):
    _fields_ = [(name, ctypes.c_void_p) for name in _FIELDS]


def _pack(parameters, class_ids, device):
    """dict of [B,Kp,(3)] tensors -> (ctypes struct of device pointers, the tensors kept alive, B, Kp)."""
    keep, st, shape = [], _Params(), None
    for cid in set(class_ids):
        for name in _NEEDS[cid]:
            if name not in parameters:
                raise KeyError("parameters lacks '%s'" % name)
            t = parameters[name].detach()
            if not t.is_cuda:
                raise RuntimeError("residues: CUDA tensors required (there is no CPU path)")
            t = t.to(torch.float32).contiguous()
            keep.append(t)
            setattr(st, name, t.data_ptr())
            if shape is None:
                shape = (t.shape[0], t.shape[1])
            elif (t.shape[0], t.shape[1]) != shape:
                raise RuntimeError("residues: parameter tensors disagree on [B, K]")
    return st, keep, shape[0], shape[1]


def _class_ids(classes):
    ids = []
    for c in classes:
        if c not in CLASS_ID:
            raise NotImplementedError
        ids.append(CLASS_ID[c])
    if not 1 <= len(ids) <= 4:
        raise RuntimeError("residues: between one and four classes")
    return ids


def residues(parameters, matching_indices, points, classes=('plane', 'sphere', 'cylinder', 'cone'), per_point=True,
             mean=True):
    """points [B,K,N',3] (a stride-0 expand over K is used as is) -> (per_point [B,K,N',T] | None,
    mean over the points [B,K,T] | None)."""
    ids = _class_ids(classes)
    dev = points.device
    st, keep, B, Kp = _pack(parameters, ids, dev)
    match = matching_indices.to(torch.int32).contiguous()
    K = match.shape[1]
    if points.dim() != 4 or points.shape[0] != B or points.shape[1] != K or points.shape[3] != 3:
        raise RuntimeError("residues: points must be [B,K,N',3]")
    if points.dtype != torch.float32:
        points = points.float()
    if points.stride(3) != 1 or points.stride(2) != 3 or (points.stride(1) not in (0, points.shape[2] * 3) and K > 1):
        points = points.contiguous()
    n_pts, T = points.shape[2], len(ids)
    out = torch.empty(B, K, n_pts, T, dtype=torch.float32, device=dev) if per_point else None
    avg = torch.empty(B, K, T, dtype=torch.float32, device=dev) if mean else None
    cls = (ctypes.c_int * T)(*ids)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().cpfn_primitive_residues(
            ctypes.byref(st), match.data_ptr(), points.data_ptr(), points.stride(0) if B > 1 else 0,
            points.stride(1) if K > 1 else 0, B, Kp, K, n_pts, cls, T, out.data_ptr() if per_point else None,
            avg.data_ptr() if mean else None, torch.cuda.current_stream(dev).cuda_stream), "primitive_residues")
    cuda_ops.count_launches(1)
    return out, avg


def p_coverage(P, prim_class, matching_indices, parameters, epsilons):
    """P [B,N,3]; prim_class int [B,K] = class id (0 plane, 1 sphere, 2 cylinder, 3 cone) of every primitive;
    epsilons: up to four floats -> coverage float32 [B, len(epsilons)]."""
    eps = [float(e) for e in epsilons]
    dev = P.device
    st, keep, B, Kp = _pack(parameters, [0, 1, 2, 3], dev)
    match = matching_indices.to(torch.int32).contiguous()
    cls = prim_class.to(torch.int32).contiguous()
    P = P.to(torch.float32).contiguous()
    K, N = match.shape[1], P.shape[1]
    count = torch.empty(B, len(eps), dtype=torch.float32, device=dev)
    eps_c = (ctypes.c_float * len(eps))(*eps)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().cpfn_p_coverage(ctypes.byref(st), match.data_ptr(), cls.data_ptr(), P.data_ptr(), B, Kp, K,
                                              N, eps_c, len(eps), count.data_ptr(),
                                              torch.cuda.current_stream(dev).cuda_stream), "p_coverage")
    cuda_ops.count_launches(1)
    return count / float(N)
