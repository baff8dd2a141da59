"""Pins oracle/network.py to the reference's own network code: tests/golden/ref_network.npz
was produced by the UNMODIFIED reference PointNet2 (tests/golden/make_ref_network_golden.py).
CPU only.  Tolerance: both sides are torch CPU fp32 with the same op sequence; 1e-5
relative to the tensor scale absorbs thread-count dependent summation order."""
import os

import numpy as np
import pytest
import torch

from cpfn_b200.pn2_network import PointNet2
from oracle import network
from tests.golden import cases

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ref_network.npz")


def _rel(a, b):
    return float(np.abs(a - b).max() / max(1e-6, np.abs(b).max()))


@pytest.fixture(scope="module")
def setup():
    model = PointNet2(output_sizes=[3, 4, 28])
    state = cases.network_state(model.state_dict())
    return {k: torch.from_numpy(v) for k, v in state.items()}, cases.network_input(), np.load(GOLDEN)


def test_state_dict_layout_matches_reference_counts():
    # 125 entries, 1 406 307 parameters for heads [3,4,28]
    m = PointNet2(output_sizes=[3, 4, 28])
    assert len(m.state_dict()) == 125
    assert sum(p.numel() for p in m.parameters()) == 1406307
    assert sum(p.numel() for p in PointNet2(output_sizes=[3, 4, 21]).parameters()) == 1405404
    assert sum(p.numel() for p in PointNet2(output_sizes=[2]).parameters()) == 1402050
    assert all(("bn" in n) for n, mod in m.named_modules() if isinstance(mod, torch.nn.modules.batchnorm._BatchNorm))


def test_network_oracle_matches_reference_golden(setup):
    sd, P, g = setup
    mask = np.unpackbits(g["mask_bits"])[: 2 * 128 * 1024].reshape(2, 128, 1024).astype(np.float32) * 2.0
    out = network.pointnet2_forward(sd, P, 3, dropout_mask=mask)
    assert _rel(out["l1_feats"][:, ::8, ::4], g["l1_feats_s"]) < 1e-5
    assert _rel(out["l2_feats"][:, ::8, ::2], g["l2_feats_s"]) < 1e-5
    assert _rel(out["l3_feats"][:, :, 0], g["l3_feats"]) < 1e-5
    assert _rel(out["l6_feats"][:, ::8, ::4], g["l6_feats_s"]) < 1e-5
    assert _rel(out["feat_pre_dropout"][:, ::8, ::4], g["feat_pre_dropout_s"]) < 1e-5
    for i in range(3):
        assert _rel(out["heads"][i], g["head%d" % i]) < 1e-5
