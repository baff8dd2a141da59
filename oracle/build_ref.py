"""Build the UNMODIFIED reference CUDA extension into oracle/_ref/.

TEST INFRASTRUCTURE.  Compiles the nine reference kernels
(/root/reference/PointNet2/pointnet2_ops/cuda_ops/src/*.{cpp,cu}) for sm_100a
from the sources where they lie -- nothing is copied into this repository and
the reference's own setup.py is not run.  The only output is
``oracle/_ref/ref_cuda_ops.so`` (git-ignored; it travels to the GPU box with
the gpurun snapshot, where /root/reference does not exist).

The module is the live parity oracle for ``pytest -m gpu`` and the source of
``tests/golden/ref_cuda_ops_*.npz`` (see tests/golden/make_ref_cuda_ops_golden.py).
It only runs on a GPU: every entry point raises "CPU not supported" otherwise.
"""
import glob
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_TREE = "/root/reference"
REF_ROOT = REF_TREE + "/PointNet2/pointnet2_ops/cuda_ops"
OUT_DIR = os.path.join(_HERE, "_ref")
NAME = "ref_cuda_ops"
# The reference's Python (unmodified files) staged for the GPU box, where /root/reference does not exist:
# baseline/_ref/ is git-ignored and travels with the gpurun snapshot.
STAGE_DIR = os.path.join(os.path.dirname(_HERE), "baseline", "_ref")
STAGED = ("PointNet2/pn2_network.py", "PointNet2/pointnet2_ops/modules", "SPFN", "Utils")


def so_path():
    return os.path.join(OUT_DIR, NAME + ".so")


def build(verbose=False):
    """Compile if the reference tree is present; return the .so path or None."""
    if os.path.exists(so_path()):
        return so_path()
    if not os.path.isdir(REF_ROOT):
        return None
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    from torch.utils.cpp_extension import load
    os.makedirs(OUT_DIR, exist_ok=True)
    sources = sorted(glob.glob(os.path.join(REF_ROOT, "src", "*.cpp")) +
                     glob.glob(os.path.join(REF_ROOT, "src", "*.cu")))
    load(name=NAME, sources=sources,
         extra_include_paths=[os.path.join(REF_ROOT, "include")],
         extra_cflags=["-O2"],
         extra_cuda_cflags=["-O2", "-gencode", "arch=compute_100a,code=sm_100a"],
         build_directory=OUT_DIR, verbose=verbose, is_python_module=True)
    return so_path() if os.path.exists(so_path()) else None


def stage_python():
    """Copy the reference's Python packages on the hot path (PointNet2 modules, SPFN, Utils) byte for byte into
    baseline/_ref/ (git-ignored).  Returns the directory, or None when neither the reference tree nor a staged
    copy exists."""
    import shutil
    if os.path.isdir(REF_TREE):
        for rel in STAGED:
            src, dst = os.path.join(REF_TREE, rel), os.path.join(STAGE_DIR, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            if os.path.isdir(src):
                shutil.copytree(src, dst, dirs_exist_ok=True, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
            else:
                shutil.copy2(src, dst)
    return STAGE_DIR if os.path.isfile(os.path.join(STAGE_DIR, "PointNet2", "pn2_network.py")) else None


def load_module():
    """Import the prebuilt extension (no compilation); None when absent."""
    path = so_path()
    if not os.path.exists(path):
        return None
    import torch  # noqa: F401  (libtorch must be loaded first)
    spec = importlib.util.spec_from_file_location(NAME, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    p = build(verbose="-v" in sys.argv)
    print("reference extension:", p)
    print("reference python staged in:", stage_python())
    if p:
        m = load_module()
        print("exports:", sorted(n for n in dir(m) if not n.startswith("_")))
