"""cpfn_b200 -- sm_100a implementation of CPFN's data-parallel hot path (PointNet++
grouping backbone + SPFN weighted-TLS fitters) behind the reference's Python API.

    cpfn_b200.cuda_ops            the nine native ops (reference pybind module `cuda_ops`)
    cpfn_b200.pointnet2_ops       PointsetAbstraction / PointsetFeaturePropagation / geometry_utils
    cpfn_b200.pn2_network         PointNet2
    cpfn_b200.spfn                fitters (plane / sphere / cylinder / cone, differentiable_tls)
    cpfn_b200.api                 GlobalSPFN engine (forward + fitters, host-buffer entry point)
    cpfn_b200.dropin.install()    registers this package under the reference's module names
"""
__version__ = "0.1.0"
