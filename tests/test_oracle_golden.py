"""Pins the C oracle to the reference: tests/golden/ref_cuda_ops.npz holds outputs
of the UNMODIFIED reference kernels (built from /root/reference into oracle/_ref,
run on a B200 by tests/golden/make_ref_cuda_ops_golden.py).  Runs on CPU."""
import os

import numpy as np
import pytest

from tests.golden import cases

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ref_cuda_ops.npz")


@pytest.fixture(scope="module")
def golden():
    assert os.path.exists(GOLDEN), "golden vectors missing"
    return np.load(GOLDEN)


def test_fps_oracle_equals_reference_kernel(oracle_ops, golden):
    for name, (xyz, m) in cases.fps_cases().items():
        np.testing.assert_array_equal(oracle_ops.farthest_point_sampling(xyz, m),
                                      golden["fps/" + name].astype(np.int32), err_msg=name)


def test_ball_query_oracle_equals_reference_kernel(oracle_ops, golden):
    for name, (q, xyz, r, k) in cases.ball_cases(oracle_ops.farthest_point_sampling).items():
        np.testing.assert_array_equal(oracle_ops.ball_query(q, xyz, r, k),
                                      golden["ball/" + name].astype(np.int32), err_msg=name)


def test_three_nn_oracle_equals_reference_kernel(oracle_ops, golden):
    for name, (u, kn) in cases.three_nn_cases(oracle_ops.farthest_point_sampling).items():
        d2, idx = oracle_ops.three_nn(u, kn)
        np.testing.assert_array_equal(idx, golden["nn_idx/" + name].astype(np.int32), err_msg=name)
        np.testing.assert_array_equal(d2, golden["nn_d2/" + name], err_msg=name)


def test_interp_gather_group_oracle_equals_reference_kernel(oracle_ops, golden):
    z = cases.interp_inputs()
    np.testing.assert_array_equal(oracle_ops.three_weighted_sum(z["pts"], z["idx"], z["w"]), golden["tws"])
    gi = np.ascontiguousarray(z["idx"][:, :, 0])
    np.testing.assert_array_equal(oracle_ops.gather_points(z["pts"], gi), golden["gather"])
    np.testing.assert_array_equal(oracle_ops.group_points(z["pts"], z["gidx"]), golden["group"])
    # scatter-adds: the reference's atomic order is nondeterministic -> tolerance
    np.testing.assert_allclose(oracle_ops.three_weighted_sum_grad(z["g"], z["idx"], z["w"], z["M"]),
                               golden["tws_grad"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(oracle_ops.gather_points_grad(z["g"], gi, z["M"]), golden["gather_grad"],
                               rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(oracle_ops.group_points_grad(z["gg"], z["gidx"], z["M"]),
                               golden["group_grad"], rtol=1e-5, atol=1e-5)
