"""oracle/sampling_lowres.py against the goldens of the unmodified reference functions (CPU)."""
import os

import numpy as np
import pytest

from oracle import sampling_lowres as olow
from tests.golden import cases

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_sampling_lowres.npz"))


@pytest.mark.parametrize("name", sorted(cases.lowres_cases()))
def test_matches_reference(name):
    P, labels, m, seed = cases.lowres_cases()[name]
    np.random.seed(seed)
    first = olow.furthest_point_sampling_per_label(P, labels)
    assert first.dtype == np.int32 and np.array_equal(first, GOLD[name + "/per_label"])
    assert len(np.unique(labels[first])) == len(np.unique(labels))       # the script's own assertion (:66)
    assert np.array_equal(olow.furthest_point_sampling(P, first, m), GOLD[name + "/fps"])
    assert np.array_equal(olow.furthest_point_sampling(P, np.zeros(0, np.int32), 64), GOLD[name + "/fps_unseeded"])
