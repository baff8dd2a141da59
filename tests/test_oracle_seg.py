"""The numpy restatement of hungarian_matching / compute_miou_loss (oracle/seg.py) against the outputs of the
unmodified reference (tests/golden/ref_seg.npz)."""
import os

import numpy as np
import pytest

from oracle import seg as oseg
from tests.golden import cases

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_seg.npz"))


@pytest.mark.parametrize("name", sorted(cases.seg_cases()))
def test_oracle_matches_reference(name):
    W, I = cases.seg_cases()[name]
    m = oseg.hungarian_matching(W, I)
    assert np.array_equal(m, GOLD[name + "/matching"])
    loss, one_minus = oseg.compute_miou_loss(W, I, m)
    assert np.abs(loss - GOLD[name + "/miou_loss"]).max() < 2e-6
    assert np.abs(one_minus - GOLD[name + "/one_minus_dot"]).max() < 2e-6
