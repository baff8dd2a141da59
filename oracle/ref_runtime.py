"""Run the UNMODIFIED reference Python (PointNet2 modules, SPFN fitters) staged under baseline/_ref/.

TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE.  oracle/build_ref.py copies the reference's own files there byte
for byte (git-ignored; they travel to the GPU box with the gpurun snapshot) and builds the reference CUDA extension
into oracle/_ref/.  This module only arranges the imports:

  * ``load_pointnet2(cuda_ops)`` imports the reference ``PointNet2.pn2_network`` with
    ``PointNet2.pointnet2_ops.cuda_ops`` bound to the given module -- the reference extension (the "before" on the
    same GPU, BASELINE.md section 3 G-ref) or ``cpfn_b200.cuda_ops`` (seam B1 of SURVEY 8b: the reference's Python on
    this package's kernels, what ``cpfn_b200.dropin.install(level="ops")`` gives a user);
  * ``load_spfn()`` imports the reference ``SPFN.losses_implementation`` with the two shims torch 2.11 needs
    (``torch.solve`` was removed; ``Tensor.get_device()`` is -1 on CPU) -- no reference file is edited.
"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGE_DIR = os.path.join(ROOT, "baseline", "_ref")
_PREFIXES = ("PointNet2", "SPFN", "Utils")


def available():
    return os.path.isfile(os.path.join(STAGE_DIR, "PointNet2", "pn2_network.py"))


def install_shims():
    if not getattr(torch.solve, "_cpfn_shim", False):           # torch 2.11 keeps a stub that raises
        def solve(B, A):
            return torch.linalg.solve(A, B), None
        solve._cpfn_shim = True
        torch.solve = solve
    if not getattr(torch.Tensor.get_device, "_cpfn_shim", False):
        orig = torch.Tensor.get_device

        def get_device(self):
            return self.device if not self.is_cuda else orig(self)
        get_device._cpfn_shim = True
        torch.Tensor.get_device = get_device


def _purge():
    for name in [n for n in sys.modules if n.split(".")[0] in _PREFIXES]:
        del sys.modules[name]


def _with_stage_on_path(fn):
    if not available():
        raise RuntimeError("baseline/_ref is not staged (run oracle/build_ref.py where /root/reference exists)")
    install_shims()
    _purge()
    sys.path.insert(0, STAGE_DIR)
    try:
        return fn()
    finally:
        sys.path.remove(STAGE_DIR)
        _purge()                      # the returned module objects stay alive; later imports start clean


def load_pointnet2(cuda_ops):
    """The reference's ``PointNet2.pn2_network`` module running on ``cuda_ops``."""
    def go():
        pkg = importlib.import_module("PointNet2.pointnet2_ops")       # namespace package of the staged tree
        sys.modules["PointNet2.pointnet2_ops.cuda_ops"] = cuda_ops
        pkg.cuda_ops = cuda_ops
        return importlib.import_module("PointNet2.pn2_network")
    return _with_stage_on_path(go)


def load_spfn():
    """The reference's ``SPFN.losses_implementation`` (and through it the four fitters)."""
    return _with_stage_on_path(lambda: importlib.import_module("SPFN.losses_implementation"))
