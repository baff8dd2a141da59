"""oracle/residues.py against the goldens of the unmodified reference (CPU), and the forwarding of the
non-hot-path names of SPFN.losses_implementation / metric_implementation to the reference's own files."""
import os
import sys

import numpy as np
import pytest

from oracle import residues as ores
from tests.golden import cases

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_residues.npz"))


def _close(a, b, tol=2e-5):
    return np.abs(a - b).max() <= tol * max(1.0, np.abs(b).max())


def test_residues_match_reference():
    params, matching, points, T_gt, P = cases.residue_case()
    loss, per_point = ores.compute_residue_loss(params, matching, points, T_gt)
    assert per_point.shape == GOLD["residue_per_point"].shape
    assert _close(per_point, GOLD["residue_per_point"]) and _close(loss, GOLD["residue_loss"])
    loss2, pp2 = ores.compute_residue_loss(params, matching, points, np.minimum(T_gt, 1), classes=('cone', 'plane'))
    assert _close(pp2, GOLD["residue_per_point_cone_plane"]) and _close(loss2, GOLD["residue_loss_cone_plane"])
    assert _close(ores.get_residual_loss(params, matching, points, T_gt), GOLD["residual"])
    for eps in (0.05, 0.2):
        got = ores.compute_P_coverage(P, T_gt, matching, params, np.float32(eps))
        assert np.abs(got - GOLD["p_coverage_%g" % eps]).max() <= 1.5 / P.shape[1]      # a point may sit on the threshold


@pytest.mark.skipif(not os.path.isdir("/root/reference/SPFN"), reason="reference checkout not present")
def test_non_hot_path_names_come_from_the_reference(built_lib):
    from cpfn_b200 import dropin
    keep = {k: v for k, v in sys.modules.items() if k.split(".")[0] in ("PointNet2", "SPFN", "Utils")}
    sys.path.insert(0, "/root/reference")
    try:
        dropin.install("full")
        from SPFN import losses_implementation, metric_implementation
        import cpfn_b200.spfn as ours
        assert losses_implementation is ours.losses_implementation
        fn = losses_implementation.compute_normal_loss                    # not restated here: the reference's own
        assert fn.__module__.startswith("cpfn_b200.spfn._ref_") and "SPFN/losses_implementation.py" in fn.__code__.co_filename
        # ... and inside the reference's module the hot-path functions are this package's
        ref = ours._reference.load("losses_implementation", {})
        assert ref.compute_parameters is ours.losses_implementation.compute_parameters
        assert ref.hungarian_matching is ours.losses_implementation.hungarian_matching        # device matching (row f3)
        assert ref.compute_miou_loss is ours.losses_implementation.compute_miou_loss
        assert ref.plane_fitter is ours.plane_fitter
        assert callable(metric_implementation.compute_Sk_coverage)
        with pytest.raises(AttributeError):
            losses_implementation.no_such_function
    finally:
        sys.path.remove("/root/reference")
        dropin.uninstall()
        for k in [k for k in sys.modules if k.startswith("cpfn_b200.spfn._ref_")]:
            del sys.modules[k]
        ours._reference._loaded.clear()
        sys.modules.update(keep)
