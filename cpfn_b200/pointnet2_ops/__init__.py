"""Mirror of the reference package ``PointNet2.pointnet2_ops``: ``cuda_ops`` (the
nine native ops, here a ctypes front-end of libcpfn_b200.so) and ``modules``."""
from .. import cuda_ops  # noqa: F401
