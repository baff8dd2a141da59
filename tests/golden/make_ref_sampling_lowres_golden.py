"""Generate tests/golden/ref_sampling_lowres.npz by running the UNMODIFIED numba functions of the reference's
Preprocessing/preprocessing_sampling_lowres.py as plain numpy (dev container only: needs /root/reference).

    python tests/golden/make_ref_sampling_lowres_golden.py

numba and h5py are not installed here: numba is stubbed (identity ``jit``, type objects that are subscriptable,
callable and carry a numpy dtype), h5py by an empty module (only the file I/O part of the script uses it).
"""
import importlib.util
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.golden import cases  # noqa: E402
from tests.golden.make_ref_merging_golden import _NumbaType  # noqa: E402


def main():
    if "numba" not in sys.modules:
        m = types.ModuleType("numba")
        m.int32, m.int64 = _NumbaType(np.int32), _NumbaType(np.int64)
        m.float32, m.float64 = _NumbaType(np.float32), _NumbaType(np.float64)
        m.jit = lambda *a, **k: (lambda f: f)
        sys.modules["numba"] = m
    if "h5py" not in sys.modules:
        sys.modules["h5py"] = types.ModuleType("h5py")
    spec = importlib.util.spec_from_file_location("ref_sampling_lowres",
                                                  "/root/reference/Preprocessing/preprocessing_sampling_lowres.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    out = {}
    for name, (P, labels, m, seed) in cases.lowres_cases().items():
        np.random.seed(seed)
        first = ref.furthest_point_sampling_per_label(P, labels)          # :28-42
        second = ref.furthest_point_sampling(P, first, m)                 # :14-26
        out[name + "/per_label"], out[name + "/fps"] = first, second
        out[name + "/fps_unseeded"] = ref.furthest_point_sampling(P, np.zeros(0, dtype=np.int32), 64)
    path = os.path.join(ROOT, "tests", "golden", "ref_sampling_lowres.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
