"""numpy restatement of the reference's patch -> object merging.

TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Follows Utils/merging_utils.py and the fusion block of
evaluation_localSPFN.py:99-130 line by line with dense numpy arrays (the literal algorithm, including the
repeated-argmax loop of heuristic_merging, so only small cases finish quickly).

Parity pin: tests/golden/ref_merging.npz, produced by running the UNMODIFIED reference module
(tests/golden/make_ref_merging_golden.py; numba is not installed here, so ``numba.jit`` is stubbed by the
identity decorator -- the jitted function is plain numpy code).
"""
import numpy as np


def point2primitive(spfn_labels, predicted_labels, point_indices, accumulate):
    """merging_utils.py:7-13 (accumulate=True, ``+=``) / evaluation_localSPFN.py:103-106 (``=``)."""
    Ng, Kg = spfn_labels.shape
    nb, Np, Kl = predicted_labels.shape
    A = np.zeros((Ng, nb * Kl + Kg), np.float32)
    for b in range(nb):
        if accumulate:
            A[point_indices[b], b * Kl:(b + 1) * Kl] += predicted_labels[b]
        else:
            A[point_indices[b], b * Kl:(b + 1) * Kl] = predicted_labels[b]
    A[:, nb * Kl:] = spfn_labels
    return A


def similarity_soft(spfn_labels, predicted_labels, point_indices, dtype=np.float32):
    """merging_utils.py:6-15.  dtype=float64 gives the exactly-rounded answer the fp32 results are judged against."""
    A = point2primitive(spfn_labels, predicted_labels, point_indices, True).astype(dtype)
    return A.T @ A


def heuristic_merging(pairs_id, patch_id, penalty_value):
    """merging_utils.py:17-33, literally."""
    pairs_id1, pairs_id2 = pairs_id[:, 0], pairs_id[:, 1]
    segment_id = np.arange(len(patch_id), dtype=np.int64)
    patch_1hot = np.eye(patch_id.max() + 1)[patch_id]
    while len(pairs_id1) > 0:
        pair_id1 = pairs_id1[np.argmax(penalty_value)]
        pair_id2 = pairs_id2[np.argmax(penalty_value)]
        segment_id[segment_id == segment_id[pair_id2]] = segment_id[pair_id1]
        selection_row = segment_id == segment_id[pair_id1]
        patch_1hot[selection_row] = np.sum(patch_1hot[selection_row], axis=0)
        intersection = np.sum(patch_1hot[pairs_id1] * patch_1hot[pairs_id2], axis=1)
        pairs_id1 = pairs_id1[intersection == 0]
        pairs_id2 = pairs_id2[intersection == 0]
        penalty_value = penalty_value[intersection == 0]
    return segment_id


def run_heuristic_solver(similarity_matrix, nb_patches, max_label_per_object, max_label_per_patch, threshold=0):
    """merging_utils.py:35-44."""
    indices = np.where(similarity_matrix > threshold)
    penalty_array = np.stack((indices[0], indices[1], similarity_matrix[indices[0], indices[1]]), axis=1)
    penalty_array = penalty_array[penalty_array[:, 0] < penalty_array[:, 1]]
    patch_id = np.concatenate((np.repeat(np.arange(nb_patches), repeats=max_label_per_patch, axis=0),
                               nb_patches * np.ones([max_label_per_object], dtype=int)), axis=0)
    labels = heuristic_merging(penalty_array[:, :2].astype(int), patch_id, penalty_array[:, 2].astype(np.float64))
    flag = np.diag(similarity_matrix)
    replacement_values = np.concatenate((np.tile(np.arange(-max_label_per_patch, 0), nb_patches),
                                         np.arange(-max_label_per_object, 0)), axis=0)
    labels[flag < threshold] = replacement_values[flag < threshold]
    _, labels = np.unique(labels, return_inverse=True)
    return labels


def get_point_final(point2primitive_prediction, output_labels_heuristic, dtype=np.float32):
    """merging_utils.py:46-50."""
    onehot = np.eye(output_labels_heuristic.max() + 1, dtype=np.float32)[output_labels_heuristic]
    onehot = onehot / (np.sum(onehot, axis=0, keepdims=True) + np.float32(1e-10))
    return point2primitive_prediction.astype(dtype) @ onehot.astype(dtype)


def fuse_patches(spfn_labels, predicted_labels, point_indices, labels, dtype=np.float32):
    """evaluation_localSPFN.py:103-111."""
    nb, Np, Kl = predicted_labels.shape
    A = point2primitive(spfn_labels, predicted_labels, point_indices, False)
    flag = np.sum(A[:, :nb * Kl], axis=1) > 0
    A[flag, nb * Kl:] = 0
    return get_point_final(A, labels, dtype)


def merge_normals_types(X, T, point_indices, spfn_normals, spfn_type):
    """evaluation_localSPFN.py:113-130 (scatter_add_ visits the flattened (patch, point) pairs in order)."""
    Ng, n_types = spfn_type.shape
    flat = point_indices.reshape(-1)
    Xg = np.zeros((Ng, 3), np.float32)
    num = np.zeros((Ng, n_types), np.float32)
    den = np.zeros((Ng,), np.float32)
    Xf, Tf = X.reshape(-1, 3).astype(np.float32), T.reshape(-1, n_types).astype(np.float32)
    for i, p in enumerate(flat):
        Xg[p] += Xf[i]
        num[p] += Tf[i]
        den[p] += np.float32(1)
    empty = np.all(Xg == 0, axis=1)
    Xg[empty] = spfn_normals[empty]
    nrm = np.sqrt(((Xg[:, 0] * Xg[:, 0] + Xg[:, 1] * Xg[:, 1]).astype(np.float32) + Xg[:, 2] * Xg[:, 2]).astype(np.float32))
    Xg = Xg / np.maximum(nrm, np.float32(1e-12))[:, None]
    Tg = num / np.maximum(den, np.float32(1))[:, None]
    Tg[empty] = spfn_type[empty]
    return Xg.astype(np.float32), Tg.astype(np.float32)
