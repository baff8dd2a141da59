"""Register cpfn_b200 under the reference's module names so that reference code runs unchanged.

    import cpfn_b200.dropin as dropin
    dropin.install()                       # before `import PointNet2...` / `import SPFN...`
    from PointNet2.pn2_network import PointNet2            # -> cpfn_b200.pn2_network.PointNet2
    from SPFN import cone_fitter, fitter_factory           # -> cpfn_b200.spfn.*

Two levels:
  level="ops"   only ``PointNet2.pointnet2_ops.cuda_ops`` is replaced (seam B1 of SURVEY.md 8b): the
                reference's own Python modules run on top of libcpfn_b200.so.  Needs the reference
                tree on sys.path.
  level="full"  the module API (B2) and the fitter API (B3) are replaced as well; the reference tree
                is not needed for the hot path.  ``Utils.sampling_utils`` (patch extraction, row f2) and
                ``Utils.merging_utils`` (patch -> object merging, row f1) are replaced too; the rest of the
                reference's ``Utils`` package is left alone.
"""
import sys
import types


def install(level="full"):
    from . import cuda_ops
    if level == "ops":
        import importlib
        pkg = importlib.import_module("PointNet2.pointnet2_ops")      # the reference package (empty __init__)
        sys.modules["PointNet2.pointnet2_ops.cuda_ops"] = cuda_ops
        pkg.cuda_ops = cuda_ops
        return
    if level != "full":
        raise ValueError("level must be 'ops' or 'full'")
    from . import pn2_network, pointnet2_ops, spfn
    from .pointnet2_ops import modules
    from .pointnet2_ops.modules import geometry_utils, pointset_abstraction, pointset_feature_propagation
    from .spfn import (cone_fitter, cylinder_fitter, fitter_factory, losses_implementation, plane_fitter,
                       sphere_fitter)
    root = types.ModuleType("PointNet2")
    root.__path__ = []
    root.pn2_network, root.pointnet2_ops = pn2_network, pointnet2_ops
    table = {
        "PointNet2": root,
        "PointNet2.pn2_network": pn2_network,
        "PointNet2.pointnet2_ops": pointnet2_ops,
        "PointNet2.pointnet2_ops.cuda_ops": cuda_ops,
        "PointNet2.pointnet2_ops.modules": modules,
        "PointNet2.pointnet2_ops.modules.geometry_utils": geometry_utils,
        "PointNet2.pointnet2_ops.modules.pointset_abstraction": pointset_abstraction,
        "PointNet2.pointnet2_ops.modules.pointset_feature_propagation": pointset_feature_propagation,
        "SPFN": spfn,
        "SPFN.plane_fitter": plane_fitter, "SPFN.sphere_fitter": sphere_fitter,
        "SPFN.cylinder_fitter": cylinder_fitter, "SPFN.cone_fitter": cone_fitter,
        "SPFN.fitter_factory": fitter_factory, "SPFN.losses_implementation": losses_implementation,
        "SPFN.differentiable_tls": spfn.differentiable_tls, "SPFN.geometry_utils": spfn.geometry_utils,
        "SPFN.metric_implementation": spfn.metric_implementation,
    }
    sys.modules.update(table)
    # Patch extraction: only the one module of the reference's ``Utils`` package is
    # replaced -- the package itself (config loader, dataset utilities, ...) stays the reference's.
    from . import merging_utils, sampling_utils
    sys.modules["Utils.sampling_utils"] = sampling_utils
    sys.modules["Utils.merging_utils"] = merging_utils              # patch -> object merging, row f1
    if "Utils" in sys.modules:
        sys.modules["Utils"].sampling_utils = sampling_utils
        sys.modules["Utils"].merging_utils = merging_utils


def uninstall():
    for name in [n for n in sys.modules if n == "PointNet2" or n.startswith("PointNet2.") or n == "SPFN"
                 or n.startswith("SPFN.") or n in ("Utils.sampling_utils", "Utils.merging_utils")]:
        mod = sys.modules[name]
        if getattr(mod, "__name__", "").startswith("cpfn_b200") or name == "PointNet2":
            del sys.modules[name]
