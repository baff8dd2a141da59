"""numpy restatement of the reference's patch extraction.

TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Same steps as Utils/sampling_utils.py:4-19 and
Preprocessing/preprocessing_sampling_patch.py:22-47 (cited per step), written around one shared helper; the
only deliberate difference is ``kind='stable'`` in the argsort (the reference's default introsort leaves the
order of EQUAL distances unspecified; stable = ordered by index, which is what the CUDA path produces).

Parity pin: tests/golden/ref_patches.npz, produced by running the UNMODIFIED reference functions in the
dev container on seeded clouds; rows compared up to the order of equal distances
(tests/golden/make_ref_patches_golden.py; h5py and numba, imported but unused by the two functions, are
stubbed at import time).
"""
import numpy as np


def nearest(seed_xyz, gt_points_hr, num_points_patch):
    """sampling_utils.py:9-13: (indices of the k nearest high-res points in distance order, their distances)."""
    d = np.linalg.norm(seed_xyz[None, :] - gt_points_hr, axis=1)
    order = np.argsort(d, kind='stable')[:num_points_patch]
    return order, d[order]


def _take_patch(lr, hr, pool, seed_index, k):
    """One patch around low-res point ``seed_index`` and the pool entries it swallows (:10-17): every pool point
    no farther from the seed than the patch's farthest member."""
    patch, patch_d = nearest(lr[seed_index], hr, k)
    to_pool = np.linalg.norm(lr[seed_index][None, :] - lr[pool], axis=1)
    return patch, np.where(to_pool <= patch_d.max())[0]


def sample(gt_points_lr, gt_points_hr, pool_indices, num_points_patch=8192, max_number_patches=32):
    """Utils/sampling_utils.py:4-19 (np.random is consumed exactly as there: one choice per patch, :8)."""
    patches, pool = [], pool_indices
    while len(patches) < max_number_patches and len(pool) != 0:
        seed_index = pool[np.random.choice(len(pool))]
        patch, swallowed = _take_patch(gt_points_lr, gt_points_hr, pool, seed_index, num_points_patch)
        patches.append(patch)
        pool = np.delete(pool, swallowed)
    return np.stack(patches, axis=0)


def sample_per_label(gt_points_lr, gt_points_hr, pool_indices, pool_labels, num_points_patch=8192,
                     max_number_patches=32):
    """Preprocessing/preprocessing_sampling_patch.py:22-47: rounds over the labels still in the pool (:26-27),
    one seed per label and round, drawn among that label's pool entries (:32-35)."""
    patches, pool, tags = [], pool_indices, pool_labels
    while len(patches) < max_number_patches and len(pool) != 0:
        for label in np.unique(tags):
            if len(patches) >= max_number_patches:
                break
            members = np.where(tags == label)[0]
            if len(members) == 0:
                continue
            seed_index = pool[np.random.choice(members)]
            patch, swallowed = _take_patch(gt_points_lr, gt_points_hr, pool, seed_index, num_points_patch)
            patches.append(patch)
            pool, tags = np.delete(pool, swallowed), np.delete(tags, swallowed)
    return np.stack(patches, axis=0)
