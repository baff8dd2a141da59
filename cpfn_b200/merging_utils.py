"""Patch -> object merging on the GPU -- mirror of the reference's ``Utils/merging_utils.py`` plus the fusion
block of ``evaluation_localSPFN.py:99-130``.

Same function names, argument order and results as the reference module:

    similarity_soft(spfn_labels, predicted_labels, point_indices)      -> [M, M] float32 (CUDA)
    heuristic_merging(pairs_id, patch_id, penalty_value)               -> int64 [n_nodes]      (host)
    run_heuristic_solver(similarity_matrix, nb_patches, max_label_per_object, max_label_per_patch, threshold=0)
    get_point_final(point2primitive_prediction, output_labels_heuristic) -> [N, L] float32 (CUDA)

The reference builds the dense point-to-primitive matrix [N_global, nb*Kl+Kg] (367 MB at 131 072 points,
2.8 GB at 1 M) and multiplies it with itself; here ``similarity_soft`` works from an inverse patch index
and never builds it (csrc/merge.cu).  ``fuse_patches`` / ``merge_normals_types`` do the evaluation script's
fusion block the same way, and ``merge_shape`` chains the whole thing.  The greedy label merge -- on the host in
numba in the reference, after a device->host copy of the similarity matrix -- is a device solve here
(``solve_labels_device``: csrc/merge_solve.cu), so that a shape's merge is a sequence of kernels with a single
4-byte read-back (the number of labels, which sizes the result); ``run_heuristic_solver`` / ``heuristic_merging``
keep the reference's host signatures (numpy in, numpy out) on a C implementation of the same pass.
"""
import ctypes

import numpy as np
import torch

from . import _lib, cuda_ops


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _f32(t):
    return t.to(torch.float32).contiguous()


def _need_cuda(*ts):
    for t in ts:
        if not t.is_cuda:
            raise RuntimeError("merging_utils: CUDA tensors required (there is no CPU path)")


def inverse_index(point_indices, num_global_points):
    """point_indices int [nb,Np] (unique inside a patch) -> int32 [nb,Ng]: position of the point in the patch or -1."""
    _need_cuda(point_indices)
    idx = point_indices.to(torch.int32).contiguous()
    nb, Np = idx.shape
    inv = torch.empty(nb, num_global_points, dtype=torch.int32, device=idx.device)
    with torch.cuda.device(idx.device):
        _lib.check(_lib.lib().cpfn_merge_inverse_index(idx.data_ptr(), nb, Np, num_global_points, inv.data_ptr(),
                                                       _stream(idx)), "merge_inverse_index")
    cuda_ops.count_launches(1)
    return idx, inv


def similarity_soft(spfn_labels, predicted_labels, point_indices, inverse=None):
    """merging_utils.py:6-15.  spfn_labels [Ng,Kg] (any dtype, used as float), predicted_labels [nb,Np,Kl],
    point_indices [nb,Np] -> intersection_primitives [nb*Kl+Kg, nb*Kl+Kg]."""
    _need_cuda(spfn_labels, predicted_labels, point_indices)
    S, W = _f32(spfn_labels), _f32(predicted_labels)
    Ng, Kg = S.shape
    nb, Np, Kl = W.shape
    idx, inv = inverse if inverse is not None else inverse_index(point_indices, Ng)
    M = nb * Kl + Kg
    out = torch.empty(M, M, dtype=torch.float32, device=W.device)
    lib = _lib.lib()
    with torch.cuda.device(W.device):
        nbytes = lib.cpfn_merge_similarity_workspace_bytes(nb, Kl, Kg)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=W.device)
        _lib.check(lib.cpfn_merge_similarity(W.data_ptr(), idx.data_ptr(), S.data_ptr(), inv.data_ptr(), nb, Np, Kl,
                                             Ng, Kg, out.data_ptr(), ws.data_ptr(), nbytes, _stream(W)),
                   "merge_similarity")
    cuda_ops.count_launches(2)
    return out


def heuristic_merging(pairs_id, patch_id, penalty_value):
    """merging_utils.py:17-33 (host).  pairs_id int64 [P,2], patch_id int64 [n], penalty_value float64 [P]."""
    pairs = np.ascontiguousarray(pairs_id, dtype=np.int64).reshape(-1, 2)
    patch = np.ascontiguousarray(patch_id, dtype=np.int64)
    pen = np.ascontiguousarray(penalty_value, dtype=np.float64)
    seg = np.empty(len(patch), dtype=np.int64)
    as_p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    _lib.check(_lib.lib().cpfn_heuristic_merging_host(as_p(pairs), as_p(pen), len(pen), as_p(patch), len(patch), as_p(seg)),
               "heuristic_merging")
    return seg


def run_heuristic_solver(similarity_matrix, nb_patches, max_label_per_object, max_label_per_patch, threshold=0):
    """merging_utils.py:35-44.  similarity_matrix: numpy [M,M]."""
    similarity_matrix = np.asarray(similarity_matrix)
    patch_id = np.concatenate((np.repeat(np.arange(nb_patches), repeats=max_label_per_patch, axis=0),
                               nb_patches * np.ones([max_label_per_object], dtype=int)), axis=0).astype(np.int64)
    if similarity_matrix.ndim != 2 or similarity_matrix.shape[0] != similarity_matrix.shape[1] or \
            similarity_matrix.shape[0] != len(patch_id):
        raise ValueError("similarity_matrix must be [nb*Kl+Kg, nb*Kl+Kg]")
    labels = np.empty(len(patch_id), dtype=np.int64)
    as_p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    if similarity_matrix.dtype == np.float32:
        # numpy compares a float32 array with a Python scalar in float32 (the scalar is cast first)
        sim = np.ascontiguousarray(similarity_matrix)
        rc = _lib.lib().cpfn_merge_solve_host_f32(as_p(sim), len(patch_id), float(np.float32(threshold)), as_p(patch_id), as_p(labels))
    else:
        sim = np.ascontiguousarray(similarity_matrix, dtype=np.float64)
        rc = _lib.lib().cpfn_merge_solve_host(as_p(sim), len(patch_id), float(threshold), as_p(patch_id), as_p(labels))
    _lib.check(rc, "merge_solve")
    flag = np.diag(similarity_matrix)
    replacement_values = np.concatenate((np.tile(np.arange(-max_label_per_patch, 0), nb_patches),
                                         np.arange(-max_label_per_object, 0)), axis=0)
    labels[flag < threshold] = replacement_values[flag < threshold]
    _, labels = np.unique(labels, return_inverse=True)
    return labels


MAX_DEVICE_PATCHES = 63          # patch sets are 64-bit masks on the device (patches + the object "patch")


def solve_labels_device(similarity_matrix, nb_patches, max_label_per_object, max_label_per_patch, threshold=0,
                        return_segments=False):
    """``run_heuristic_solver`` (merging_utils.py:35-44, incl. heuristic_merging :17-33) on the device, from the
    float32 CUDA matrix ``similarity_soft`` returns.  -> (labels int32 [M] CUDA, label_weight f32 [M] CUDA with
    entry l = 1 / (members of label l + 1e-10), n_labels int32 [1] CUDA[, raw segment ids int32 [M]]).  Nothing is
    copied to the host and nothing synchronises."""
    _need_cuda(similarity_matrix)
    nb, Kg, Kl = int(nb_patches), int(max_label_per_object), int(max_label_per_patch)
    M = nb * Kl + Kg
    if similarity_matrix.dtype != torch.float32 or tuple(similarity_matrix.shape) != (M, M):
        raise ValueError("similarity_matrix must be float32 [nb*Kl+Kg, nb*Kl+Kg]")
    if nb > MAX_DEVICE_PATCHES:
        raise ValueError("solve_labels_device handles up to %d patches (use run_heuristic_solver)" % MAX_DEVICE_PATCHES)
    sim = similarity_matrix.contiguous()
    dev = sim.device
    labels = torch.empty(M, dtype=torch.int32, device=dev)
    weights = torch.empty(M, dtype=torch.float32, device=dev)
    n_labels = torch.empty(1, dtype=torch.int32, device=dev)
    segments = torch.empty(M, dtype=torch.int32, device=dev) if return_segments else None
    lib = _lib.lib()
    with torch.cuda.device(dev):
        nbytes = lib.cpfn_merge_solve_workspace_bytes(nb, Kl, Kg)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        _lib.check(lib.cpfn_merge_solve(sim.data_ptr(), nb, Kl, Kg, float(np.float32(threshold)), labels.data_ptr(),
                                        weights.data_ptr(), n_labels.data_ptr(),
                                        segments.data_ptr() if segments is not None else None, ws.data_ptr(), nbytes,
                                        _stream(sim)), "merge_solve")
    cuda_ops.count_launches(4)
    return (labels, weights, n_labels, segments) if return_segments else (labels, weights, n_labels)


def _labels_and_weights(labels, device):
    """labels (numpy / list / tensor) -> (int32 labels on the device, 1/(members + 1e-10) per label, L).  The label
    vector has a few hundred entries and normally comes from the host solver: counted on the host, no sync."""
    host = labels.detach().cpu().numpy() if torch.is_tensor(labels) else np.asarray(labels)
    host = host.astype(np.int64).reshape(-1)
    L = int(host.max()) + 1
    counts = np.bincount(host, minlength=L).astype(np.float32)
    weights = (np.float32(1) / (counts + np.float32(1e-10))).astype(np.float32)      # one_hot / (sum + 1e-10), :48
    return (torch.from_numpy(host.astype(np.int32)).to(device), torch.from_numpy(weights).to(device), L)


def get_point_final(point2primitive_prediction, output_labels_heuristic):
    """merging_utils.py:46-50 on the reference's own (dense) operands."""
    _need_cuda(point2primitive_prediction)
    A = _f32(point2primitive_prediction)
    Ng, M = A.shape
    labels, v, L = _labels_and_weights(output_labels_heuristic, A.device)
    out = torch.empty(Ng, L, dtype=torch.float32, device=A.device)
    with torch.cuda.device(A.device):
        _lib.check(_lib.lib().cpfn_merge_dense_labels(A.data_ptr(), labels.data_ptr(), v.data_ptr(), Ng, M, L,
                                                      out.data_ptr(), _stream(A)), "merge_dense_labels")
    cuda_ops.count_launches(1)
    return out


def fuse_patches(spfn_labels, predicted_labels, point_indices, labels, inverse=None, device_solution=None):
    """evaluation_localSPFN.py:103-111 in one kernel: what ``get_point_final(point2primitive_fusion, labels)``
    returns there, without building point2primitive_fusion.  ``device_solution`` = (labels int32 CUDA, label_weight,
    number of labels as a Python int) from ``solve_labels_device`` skips the host-side label bookkeeping."""
    _need_cuda(spfn_labels, predicted_labels, point_indices)
    S, W = _f32(spfn_labels), _f32(predicted_labels)
    Ng, Kg = S.shape
    nb, Np, Kl = W.shape
    idx, inv = inverse if inverse is not None else inverse_index(point_indices, Ng)
    labels, v, L = device_solution if device_solution is not None else _labels_and_weights(labels, W.device)
    out = torch.empty(Ng, L, dtype=torch.float32, device=W.device)
    with torch.cuda.device(W.device):
        _lib.check(_lib.lib().cpfn_merge_point_labels(W.data_ptr(), S.data_ptr(), inv.data_ptr(), labels.data_ptr(),
                                                      v.data_ptr(), nb, Np, Kl, Ng, Kg, L, out.data_ptr(), _stream(W)),
                   "merge_point_labels")
    cuda_ops.count_launches(1)
    return out


def merge_normals_types(X, T, point_indices, spfn_normals, spfn_type, inverse=None):
    """evaluation_localSPFN.py:113-130.  X [nb,Np,3], T [nb,Np,n_types] -> (X_global [Ng,3], T_global [Ng,n_types])."""
    _need_cuda(X, T, point_indices, spfn_normals, spfn_type)
    X, T, on, ot = _f32(X), _f32(T), _f32(spfn_normals), _f32(spfn_type)
    nb, Np, _ = X.shape
    Ng, n_types = ot.shape
    idx, inv = inverse if inverse is not None else inverse_index(point_indices, Ng)
    Xg = torch.empty(Ng, 3, dtype=torch.float32, device=X.device)
    Tg = torch.empty(Ng, n_types, dtype=torch.float32, device=X.device)
    with torch.cuda.device(X.device):
        _lib.check(_lib.lib().cpfn_merge_normals_types(X.data_ptr(), T.data_ptr(), inv.data_ptr(), on.data_ptr(),
                                                       ot.data_ptr(), nb, Np, Ng, n_types, Xg.data_ptr(), Tg.data_ptr(),
                                                       _stream(X)), "merge_normals_types")
    cuda_ops.count_launches(1)
    return Xg, Tg


def merge_shape(W, X, T, point_indices, spfn_labels, spfn_normals, spfn_type, threshold=0, solver="device"):
    """The whole fusion block of evaluation_localSPFN.py:99-130 for one shape.  W [nb,Np,Kl] soft-maxed
    memberships, X [nb,Np,3] unit normals, T [nb,Np,n_types].  Returns (W_fusion [Ng,L], X_global, T_global,
    labels int64 [nb*Kl+Kg] numpy).  ``solver``: "device" (csrc/merge_solve.cu; the one host read-back is the
    number of labels) or "host" (the reference's arrangement: matrix to the host, C greedy pass there)."""
    nb, Np, Kl = W.shape
    Ng, Kg = spfn_labels.shape
    inverse = inverse_index(point_indices, Ng)
    sim = similarity_soft(spfn_labels, W, point_indices, inverse=inverse)
    if solver == "device" and nb <= MAX_DEVICE_PATCHES:
        labels_dev, weights, n_labels = solve_labels_device(sim, nb, Kg, Kl, threshold=threshold)
        X_global, T_global = merge_normals_types(X, T, point_indices, spfn_normals, spfn_type, inverse=inverse)
        L = int(n_labels.item())                                        # sizes W_fusion: the only synchronisation
        W_fusion = fuse_patches(spfn_labels, W, point_indices, None, inverse=inverse,
                                device_solution=(labels_dev, weights, L))
        return W_fusion, X_global, T_global, labels_dev.cpu().numpy().astype(np.int64)
    labels = run_heuristic_solver(sim.cpu().numpy(), nb, Kg, Kl, threshold=threshold)      # host, as in the reference
    W_fusion = fuse_patches(spfn_labels, W, point_indices, labels, inverse=inverse)
    X_global, T_global = merge_normals_types(X, T, point_indices, spfn_normals, spfn_type, inverse=inverse)
    return W_fusion, X_global, T_global, labels
