"""oracle/patches.py against the goldens produced by the unmodified reference (CPU)."""
import os

import numpy as np
import pytest

from oracle import patches as opatch
from tests.golden import cases

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_patches.npz"))


@pytest.mark.parametrize("name", sorted(cases.patch_cases()))
def test_sample_matches_reference(name):
    lr, hr, pool, labels, k, mp, seed = cases.patch_cases()[name]
    np.random.seed(seed)
    got = opatch.sample(lr, hr, pool.copy(), num_points_patch=k, max_number_patches=mp)
    assert got.dtype == np.int64
    cases.assert_patches_equivalent(got, GOLD[name + "/sample"], hr)
    np.random.seed(seed)
    got = opatch.sample_per_label(lr, hr, pool.copy(), labels.copy(), num_points_patch=k, max_number_patches=mp)
    cases.assert_patches_equivalent(got, GOLD[name + "/sample_per_label"], hr)


def test_distance_arithmetic_is_plain_fp32():
    """np.linalg.norm over float32 rows of 3 = sqrt((dx*dx + dy*dy) + dz*dz), each step rounded to fp32 --
    the sequence csrc/patch_select.cu pins with _rn intrinsics."""
    lr, hr, pool, _, k, _, _ = cases.patch_cases()["small"]
    s = lr[pool[0]]
    d = hr - s[None]                      # sign is irrelevant after squaring
    sq = (d * d).astype(np.float32)
    manual = np.sqrt(((sq[:, 0] + sq[:, 1]).astype(np.float32) + sq[:, 2]).astype(np.float32))
    ref = np.linalg.norm(s[None] - hr, axis=1)
    assert ref.dtype == np.float32 and np.array_equal(manual, ref)
    assert np.array_equal(np.sort(ref)[:k], GOLD["small/dist0"])


def test_ties_are_ordered_by_index():
    hr = np.zeros((64, 3), np.float32)
    hr[:, 0] = np.repeat(np.arange(8), 8)            # eight groups of eight identical points
    idx, dist = opatch.nearest(np.zeros(3, np.float32), hr, 12)
    assert idx.tolist() == list(range(12)) and dist.tolist() == [0.0] * 8 + [1.0] * 4


def test_host_mirror_rejects_float64_clouds(built_lib):
    """numpy computes the reference's distances in the wider dtype of the two clouds; only the float32 arithmetic
    is implemented, so float64 inputs are refused before any GPU work (no silent down-cast)."""
    from cpfn_b200 import sampling_utils
    pts = np.random.RandomState(0).randn(50, 3)
    with pytest.raises(TypeError):
        sampling_utils.sample(pts, pts.astype(np.float32), np.arange(10), 16, 2)
    with pytest.raises(TypeError):
        sampling_utils.sample_per_label(pts, pts.astype(np.float32), np.arange(10), np.zeros(10, int), 16, 2)
