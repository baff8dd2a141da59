"""One C-ABI call (cpfn_fit_primitives, include/cpfn_b200.h) fits all four primitive types
for every (cloud, instance slot): the fused replacement of
``SPFN/losses_implementation.py:255-278`` and the four ``compute_parameters`` it calls.
Under autograd (W or X requires grad) the differentiable path of ``spfn/_train.py`` is used.
No CPU fallback: CPU tensors raise."""
import torch

from .. import _lib, cuda_ops

KEYS = (("plane_normal", 0, 3), ("plane_center", 3, 1), ("sphere_center", 4, 3),
        ("sphere_radius_squared", 7, 1), ("cylinder_axis", 8, 3), ("cylinder_center", 11, 3),
        ("cylinder_radius_squared", 14, 1), ("cone_apex", 15, 3), ("cone_axis", 18, 3),
        ("cone_half_angle", 21, 1))

def _workspace(nbytes, device):
    """Scratch for the partial sums: a fresh block from torch's caching allocator per call, so that concurrent
    callers on different streams never share it and a captured CUDA graph owns the block it was captured with."""
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def fit_primitives(P, W, X):
    """P [B,N,3], W [B,N,K], X [B,N,3] float32 CUDA -> dict of the ten reference keys
    (each a contiguous view of one [22*B*K] buffer)."""
    return fit_primitives_packed(P, W, X)[0]


def fit_primitives_packed(P, W, X):
    """As ``fit_primitives`` but also returns the single [22*B*K] buffer the ten tensors are views of
    (one device-to-host copy moves all parameters); ``None`` on the differentiable path."""
    if not P.is_cuda:
        raise RuntimeError("CPU not supported")
    if torch.is_grad_enabled() and (W.requires_grad or X.requires_grad):
        # training: CUDA moment kernels with a backward + float64 autograd algebra (spfn/_train.py)
        from . import _train
        return _train.compute_parameters(P, W, X, ("plane", "sphere", "cylinder", "cone")), None
    B, N, _ = P.shape
    K = W.shape[2]
    P = P.detach().float().contiguous()
    W = W.detach().float().contiguous()
    X = X.detach().float().contiguous()
    out = torch.empty(22 * B * K, dtype=torch.float32, device=P.device)
    L = _lib.lib()
    with torch.cuda.device(P.device):
        nbytes = L.cpfn_fit_workspace_bytes(B, N, K)
        ws = _workspace(nbytes, P.device)
        _lib.check(L.cpfn_fit_primitives(P.data_ptr(), W.data_ptr(), X.data_ptr(), B, N, K,
                                         out.data_ptr(), ws.data_ptr(), ws.numel(),
                                         torch.cuda.current_stream(P.device).cuda_stream),
                   "fit_primitives")
    cuda_ops.count_launches(4)
    BK = B * K
    res = {}
    for name, off, width in KEYS:
        seg = out.narrow(0, off * BK, width * BK)
        res[name] = seg.view(B, K, 3) if width == 3 else seg.view(B, K)
    return res, out
