"""Patch extraction on the GPU -- host-side mirror of the reference's ``Utils/sampling_utils.py`` and of
``sample`` in ``Preprocessing/preprocessing_sampling_patch.py:22-47``.

``sample(gt_points_lr, gt_points_hr, pool_indices, num_points_patch, max_number_patches)`` keeps the
reference's signature, its use of ``np.random`` (one ``np.random.choice`` per patch, so a seeded run
picks the same seeds) and its result (int64 [n_patches, num_points_patch], each row ordered by distance).
The O(N_hr) work per seed -- distances to every high-resolution point and the selection of the nearest
``num_points_patch`` -- runs in ``cpfn_extract_patches`` (csrc/patch_select.cu); the seed choice and the
pruning of the <= 8192-entry pool stay on the host, as they are sequential and tiny.  The high-resolution
cloud is uploaded once per call and stays resident.  Equal distances are ordered by index (the reference's
numpy introsort leaves them unordered).
"""
import numpy as np
import torch

from . import _lib, cuda_ops

def _workspace(dev, nbytes):
    """Per-call scratch from torch's caching allocator (stream- and graph-safe)."""
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=dev)


def extract_patches(points_hr, seeds, num_points_patch=8192, return_distances=False):
    """points_hr f32 [N,3], seeds f32 [S,3] (CUDA tensors) -> patch indices int32 [S,k] ordered by
    (distance, index), k = min(num_points_patch, N); with ``return_distances`` also the sorted distances
    f32 [S,k] and the patch radii f32 [S].  All seeds are processed by the same launches."""
    if not (points_hr.is_cuda and seeds.is_cuda):
        raise RuntimeError("extract_patches: CUDA tensors required (there is no CPU path)")
    if points_hr.dtype != torch.float32 or seeds.dtype != torch.float32:
        raise RuntimeError("extract_patches: float32 required (the reference's data are float32)")
    if points_hr.dim() != 2 or points_hr.shape[1] != 3 or seeds.dim() != 2 or seeds.shape[1] != 3:
        raise RuntimeError("extract_patches: expected [N,3] points and [S,3] seeds")
    points_hr, seeds = points_hr.contiguous(), seeds.contiguous()
    dev = points_hr.device
    N, S = points_hr.shape[0], seeds.shape[0]
    k = min(int(num_points_patch), N)
    idx = torch.empty(S, k, dtype=torch.int32, device=dev)
    dist = torch.empty(S, k, dtype=torch.float32, device=dev) if return_distances else None
    radius = torch.empty(S, dtype=torch.float32, device=dev)
    lib = _lib.lib()
    with torch.cuda.device(dev):
        nbytes = lib.cpfn_extract_patches_workspace_bytes(N, S, k)
        ws = _workspace(dev, nbytes)
        _lib.check(lib.cpfn_extract_patches(points_hr.data_ptr(), N, seeds.data_ptr(), S, k, idx.data_ptr(),
                                            dist.data_ptr() if dist is not None else None, radius.data_ptr(),
                                            ws.data_ptr(), ws.numel(), torch.cuda.current_stream(dev).cuda_stream),
                   "extract_patches")
    cuda_ops.count_launches(N_LAUNCHES(N))
    if return_distances:
        return idx, dist, radius
    return idx


def N_LAUNCHES(N):
    """Kernels cpfn_extract_patches launches for an N-point cloud (init, 3 distance digits, index digits,
    compaction, sort)."""
    nb = max(1, int(N - 1).bit_length())
    return 1 + 3 + (1 if nb <= 11 else 2 if nb <= 22 else 3) + 2


class _Extractor:
    """High-resolution cloud resident on the device + pinned one-seed staging; one call per seed."""

    def __init__(self, gt_points_hr, num_points_patch, device):
        hr = np.ascontiguousarray(gt_points_hr)
        if hr.dtype != np.float32:
            raise TypeError("gt_points_hr must be float32 (numpy computes the reference's distances in the "
                            "array's own dtype; only the float32 arithmetic is implemented), got %s" % hr.dtype)
        self.dev = torch.device(device)
        self.hr = torch.from_numpy(hr).to(self.dev)
        self.k = min(int(num_points_patch), hr.shape[0])
        self.seed_host = torch.empty(1, 3, dtype=torch.float32).pin_memory()
        self.seed_dev = torch.empty(1, 3, dtype=torch.float32, device=self.dev)

    def __call__(self, seed_xyz):
        self.seed_host[0] = torch.from_numpy(np.asarray(seed_xyz, dtype=np.float32))
        self.seed_dev.copy_(self.seed_host, non_blocking=True)
        idx, dist, radius = extract_patches(self.hr, self.seed_dev, self.k, return_distances=True)
        radius_host = float(radius.item())                      # the pool pruning needs it now (one 4-byte read)
        return idx[0], radius_host


def _check_lr(gt_points_lr):
    """numpy would compute ``seed - gt_points_hr`` in the wider of the two dtypes: with a float64 low-res cloud the
    reference's distances are float64 arithmetic, which the float32 kernel does not reproduce."""
    if gt_points_lr.dtype != np.float32:
        raise TypeError("gt_points_lr must be float32 (as the reference's preprocessing writes it), got %s"
                        % gt_points_lr.dtype)


def _prune(gt_points_lr, i, pool_indices, radius):
    distances = np.linalg.norm(np.expand_dims(gt_points_lr[i], axis=0) - gt_points_lr[pool_indices], axis=1)
    return np.where(distances <= radius)[0]


def sample(gt_points_lr, gt_points_hr, pool_indices, num_points_patch=8192, max_number_patches=32, device="cuda:0"):
    """Utils/sampling_utils.py:4-19.  Returns int64 [n_patches, num_points_patch]."""
    gt_points_lr = np.asarray(gt_points_lr)
    _check_lr(gt_points_lr)
    pool_indices = np.asarray(pool_indices)
    extractor = _Extractor(gt_points_hr, num_points_patch, device)
    patches = []
    while (len(patches) < max_number_patches) and (len(pool_indices) != 0):
        i = pool_indices[np.random.choice(len(pool_indices))]
        idx, radius = extractor(gt_points_lr[i])
        patches.append(idx)
        pool_indices = np.delete(pool_indices, _prune(gt_points_lr, i, pool_indices, np.float32(radius)))
    return torch.stack(patches, dim=0).to(torch.int64).cpu().numpy()


def sample_per_label(gt_points_lr, gt_points_hr, pool_indices, pool_labels, num_points_patch=8192,
                     max_number_patches=32, device="cuda:0"):
    """``sample`` of Preprocessing/preprocessing_sampling_patch.py:22-47: round-robin over the labels that
    still have pool points, one seed per label and round."""
    gt_points_lr = np.asarray(gt_points_lr)
    _check_lr(gt_points_lr)
    pool_indices, pool_labels = np.asarray(pool_indices), np.asarray(pool_labels)
    extractor = _Extractor(gt_points_hr, num_points_patch, device)
    patches = []
    while (len(patches) < max_number_patches) and (len(pool_indices) != 0):
        for label in np.unique(pool_labels):
            if len(patches) >= max_number_patches:
                break
            members = np.where(pool_labels == label)[0]
            if len(members) == 0:
                continue
            i = pool_indices[np.random.choice(members)]
            idx, radius = extractor(gt_points_lr[i])
            patches.append(idx)
            gone = _prune(gt_points_lr, i, pool_indices, np.float32(radius))
            pool_indices = np.delete(pool_indices, gone)
            pool_labels = np.delete(pool_labels, gone)
    return torch.stack(patches, dim=0).to(torch.int64).cpu().numpy()
