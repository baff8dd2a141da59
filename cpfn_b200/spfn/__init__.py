"""SPFN fitter interface (reference package ``SPFN``) on the sm_100a kernels of
libcpfn_b200.so: plane_fitter / sphere_fitter / cylinder_fitter / cone_fitter
``compute_parameters``, ``losses_implementation.compute_parameters`` and ``fitter_factory``."""
from . import fit, fitter_factory, losses_implementation  # noqa: F401
from . import plane_fitter, sphere_fitter, cylinder_fitter, cone_fitter  # noqa: F401
from . import differentiable_tls, geometry_utils  # noqa: F401
from . import metric_implementation, residues, seg  # noqa: F401
