"""Low-resolution sub-sampling on the GPU -- mirror of the two numba functions of the reference's
``Preprocessing/preprocessing_sampling_lowres.py:14-42``:

    furthest_point_sampling(input_points, index_query_points1, nb_query_points)  -> int32 [nb_query_points]
    furthest_point_sampling_per_label(input_points, labels)                      -> int32 [n_labels]

Same arguments, same results (true distances, first-maximum tie-break, seeds that only start at distance 0, one
sample per label with ``np.random.randint`` choosing the start) -- bit-exact with the reference's code run as plain
numpy.  The O(N x samples) work runs in ``cpfn_fps_dense`` (csrc/fps_dense.cu): the whole GPU on one cloud, points
and running minima in registers.  ``lowres_indices`` is the composition the script applies per shape (:63-65)."""
import numpy as np
import torch

from . import _lib, cuda_ops


def _run(points, labels, seeds, start_index, n_out, device):
    pts = np.ascontiguousarray(points)
    if pts.dtype != np.float32 or pts.ndim != 2 or pts.shape[1] != 3:
        raise TypeError("input_points must be float32 [N,3] (the reference's numba signature), got %s %s"
                        % (pts.dtype, pts.shape))
    dev = torch.device(device)
    P = torch.from_numpy(pts).to(dev)
    L = torch.from_numpy(np.ascontiguousarray(labels, dtype=np.int32)).to(dev) if labels is not None else None
    S = torch.from_numpy(np.ascontiguousarray(seeds, dtype=np.int32)).to(dev) if seeds is not None and len(seeds) else None
    out = torch.empty(n_out, dtype=torch.int32, device=dev)
    lib = _lib.lib()
    with torch.cuda.device(dev):
        ws = torch.empty(lib.cpfn_fps_dense_workspace_bytes(), dtype=torch.uint8, device=dev)
        _lib.check(lib.cpfn_fps_dense(P.data_ptr(), P.shape[0], L.data_ptr() if L is not None else None,
                                      S.data_ptr() if S is not None else None, 0 if S is None else S.numel(),
                                      int(start_index), int(n_out), out.data_ptr(), ws.data_ptr(), ws.numel(),
                                      torch.cuda.current_stream(dev).cuda_stream), "fps_dense")
    cuda_ops.count_launches(1)
    return out.cpu().numpy()


def furthest_point_sampling(input_points, index_query_points1, nb_query_points, device="cuda:0"):
    """preprocessing_sampling_lowres.py:14-26."""
    if nb_query_points <= 0:
        return np.zeros(0, dtype=np.int32)
    return _run(input_points, None, index_query_points1, 0, nb_query_points, device)


def furthest_point_sampling_per_label(input_points, labels, device="cuda:0"):
    """preprocessing_sampling_lowres.py:28-42 (one ``np.random.randint`` draw, as there)."""
    num_points = len(input_points)
    n_labels = len(np.unique(labels))
    index = np.random.randint(0, num_points)
    return _run(input_points, labels, None, index, n_labels, device)


def lowres_indices(gt_points, gt_labels, nb_query_points=8192, device="cuda:0"):
    """:63-65 -- one point per label, then nb_query_points farthest points seeded with them."""
    first = furthest_point_sampling_per_label(gt_points, gt_labels, device=device)
    second = furthest_point_sampling(gt_points, first, nb_query_points, device=device)
    return np.concatenate((first, second))
