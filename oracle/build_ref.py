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
REF_ROOT = "/root/reference/PointNet2/pointnet2_ops/cuda_ops"
OUT_DIR = os.path.join(_HERE, "_ref")
NAME = "ref_cuda_ops"


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
    if p:
        m = load_module()
        print("exports:", sorted(n for n in dir(m) if not n.startswith("_")))
