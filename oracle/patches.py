"""numpy restatement of the reference's patch extraction.

TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Follows Utils/sampling_utils.py:4-19 and
Preprocessing/preprocessing_sampling_patch.py:22-47 line by line; the only deliberate difference is
``kind='stable'`` in the argsort (the reference's default introsort leaves the order of EQUAL distances
unspecified; stable = ordered by index, which is what the CUDA path produces).

Parity pin: tests/golden/ref_patches.npz, produced by running the UNMODIFIED reference functions in the
dev container on seeded clouds; rows compared up to the order of equal distances (tests/golden/make_ref_patches_golden.py; h5py and
numba, imported but unused by the two functions, are stubbed at import time).
"""
import numpy as np


def nearest(seed_xyz, gt_points_hr, num_points_patch):
    """sampling_utils.py:9-13.  (patch indices int64 [k], sorted patch distances [k])."""
    distances = np.linalg.norm(np.expand_dims(seed_xyz, axis=0) - gt_points_hr, axis=1)
    order = np.argsort(distances, kind='stable')[:num_points_patch]
    return order, distances[order]


def sample(gt_points_lr, gt_points_hr, pool_indices, num_points_patch=8192, max_number_patches=32):
    """Utils/sampling_utils.py:4-19 (np.random is consumed exactly as there)."""
    list_patch_indices = []
    while (len(list_patch_indices) < max_number_patches) and (len(pool_indices) != 0):
        i = pool_indices[np.random.choice(len(pool_indices))]                                   # :8
        patch_indices, patch_distances = nearest(gt_points_lr[i], gt_points_hr, num_points_patch)   # :10-13
        list_patch_indices.append(patch_indices)
        distances = np.linalg.norm(np.expand_dims(gt_points_lr[i], axis=0) - gt_points_lr[pool_indices], axis=1)  # :15
        pool_indices_selected = np.where(distances <= np.max(patch_distances))[0]               # :16
        pool_indices = np.delete(pool_indices, pool_indices_selected)                           # :17
    return np.stack(list_patch_indices, axis=0)


def sample_per_label(gt_points_lr, gt_points_hr, pool_indices, pool_labels, num_points_patch=8192,
                     max_number_patches=32):
    """Preprocessing/preprocessing_sampling_patch.py:22-47."""
    list_patch_indices = []
    while (len(list_patch_indices) < max_number_patches) and (len(pool_indices) != 0):
        for label in np.unique(pool_labels):                                                    # :26-27
            if len(list_patch_indices) >= max_number_patches:
                break
            ind_pool_indices = np.where(pool_labels == label)[0]                                # :32
            if len(ind_pool_indices) == 0:
                continue
            i = pool_indices[np.random.choice(ind_pool_indices)]                                # :35
            patch_indices, patch_distances = nearest(gt_points_lr[i], gt_points_hr, num_points_patch)
            list_patch_indices.append(patch_indices)
            distances = np.linalg.norm(np.expand_dims(gt_points_lr[i], axis=0) - gt_points_lr[pool_indices], axis=1)
            pool_indices_selected = np.where(distances <= np.max(patch_distances))[0]
            pool_indices = np.delete(pool_indices, pool_indices_selected)
            pool_labels = np.delete(pool_labels, pool_indices_selected)
    return np.stack(list_patch_indices, axis=0)
