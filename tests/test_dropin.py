"""The drop-in registration maps the reference's module names onto this package (CPU, no compute)."""
import sys

import pytest


def test_full_install_maps_reference_names(built_lib):
    from cpfn_b200 import dropin
    saved = {k: v for k, v in sys.modules.items() if k == "PointNet2" or k.startswith("PointNet2.") or k == "SPFN" or k.startswith("SPFN.")}
    try:
        dropin.install("full")
        from PointNet2.pn2_network import PointNet2
        from PointNet2.pointnet2_ops import cuda_ops
        from PointNet2.pointnet2_ops.modules.pointset_abstraction import PointsetAbstraction
        from PointNet2.pointnet2_ops.modules.geometry_utils import farthest_point_sample, ball_query, three_nn  # noqa: F401
        from SPFN import cone_fitter, fitter_factory, losses_implementation
        import cpfn_b200
        assert PointNet2.__module__.startswith("cpfn_b200")
        assert PointsetAbstraction.__module__.startswith("cpfn_b200")
        assert cuda_ops is cpfn_b200.cuda_ops
        for name in ("farthest_point_sampling", "ball_query", "gather_points", "gather_points_grad", "group_points",
                     "group_points_grad", "three_nn", "three_weighted_sum", "three_weighted_sum_grad"):
            assert callable(getattr(cuda_ops, name))          # bindings.cpp:7-18 of the reference
        fitter_factory.register_primitives(["sphere", "plane", "cylinder", "cone"])
        assert fitter_factory.primitive_name_to_id("plane") == 1 and fitter_factory.get_n_registered_primitives() == 4
        assert callable(cone_fitter.compute_parameters) and callable(losses_implementation.compute_parameters)
        import inspect
        mu = sys.modules["Utils.merging_utils"]                 # evaluation_localSPFN.py:101-102,111
        assert mu is cpfn_b200.merging_utils
        for fn, args in (("similarity_soft", ["spfn_labels", "predicted_labels", "point_indices"]),
                         ("heuristic_merging", ["pairs_id", "patch_id", "penalty_value"]),
                         ("run_heuristic_solver", ["similarity_matrix", "nb_patches", "max_label_per_object",
                                                   "max_label_per_patch", "threshold"]),
                         ("get_point_final", ["point2primitive_prediction", "output_labels_heuristic"])):
            assert list(inspect.signature(getattr(mu, fn)).parameters)[:len(args)] == args, fn
        su = sys.modules["Utils.sampling_utils"]                # evaluation_PatchSelection.py:14,87
        assert su is cpfn_b200.sampling_utils
        assert list(inspect.signature(su.sample).parameters)[:5] == ["gt_points_lr", "gt_points_hr", "pool_indices",
                                                                     "num_points_patch", "max_number_patches"]
    finally:
        dropin.uninstall()
        sys.modules.update(saved)


def test_cpu_tensors_are_rejected_like_the_reference(built_lib):
    import torch
    from cpfn_b200 import cuda_ops
    with pytest.raises(RuntimeError, match="CPU not supported"):
        cuda_ops.farthest_point_sampling(torch.zeros(1, 8, 3), 2)
    with pytest.raises(RuntimeError, match="float"):
        cuda_ops.farthest_point_sampling(torch.zeros(1, 8, 3, dtype=torch.float64), 2)
    with pytest.raises(RuntimeError, match="contiguous"):
        cuda_ops.ball_query(torch.zeros(1, 3, 4).transpose(1, 2), torch.zeros(1, 8, 3), 0.1, 4)
