"""Generate tests/golden/ref_patches.npz by running the UNMODIFIED reference patch-extraction functions
(dev container only: needs /root/reference).

    python tests/golden/make_ref_patches_golden.py

``Utils/sampling_utils.sample`` is pure numpy.  ``Preprocessing/preprocessing_sampling_patch.py`` imports
h5py and numba at module level (neither is installed here, neither is used by ``sample``): empty stub
modules are registered for the import; no reference file is edited.
"""
import importlib.util
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.golden import cases  # noqa: E402


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    for missing in ("h5py", "numba"):
        if missing not in sys.modules:
            try:
                __import__(missing)
            except ImportError:
                sys.modules[missing] = types.ModuleType(missing)
    ref_a = load("/root/reference/Utils/sampling_utils.py", "ref_sampling_utils")
    ref_b = load("/root/reference/Preprocessing/preprocessing_sampling_patch.py", "ref_sampling_patch")
    out = {}
    for name, (lr, hr, pool, labels, k, mp, seed) in cases.patch_cases().items():
        np.random.seed(seed)
        out[name + "/sample"] = ref_a.sample(lr, hr, pool.copy(), num_points_patch=k, max_number_patches=mp)
        np.random.seed(seed)
        out[name + "/sample_per_label"] = ref_b.sample(lr, hr, pool.copy(), labels.copy(), num_points_patch=k,
                                                       max_number_patches=mp)
        d = np.linalg.norm(lr[pool[0]][None] - hr, axis=1)
        out[name + "/dist0"] = np.sort(d)[:k]
    path = os.path.join(ROOT, "tests", "golden", "ref_patches.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
