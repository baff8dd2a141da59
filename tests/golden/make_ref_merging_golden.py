"""Generate tests/golden/ref_merging.npz by running the UNMODIFIED reference Utils/merging_utils.py on CPU
tensors (dev container only: needs /root/reference).

    python tests/golden/make_ref_merging_golden.py

numba is not installed here.  ``heuristic_merging`` is decorated with ``numba.jit(signature, nopython=True)``
and uses ``numba.int64`` as a numpy dtype; a stub module provides an identity ``jit`` and type objects that
are subscriptable / callable (to build the signature) and carry a ``dtype`` (so numpy accepts them).  The
function body then runs as the plain numpy code it is.  No reference file is edited.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.golden import cases  # noqa: E402


class _NumbaType:
    def __init__(self, dtype):
        self.dtype = np.dtype(dtype)

    def __getitem__(self, item):
        return self

    def __call__(self, *args):
        return self


def stub_numba():
    try:
        import numba  # noqa: F401
        return
    except ImportError:
        pass
    m = types.ModuleType("numba")
    m.int64, m.float64 = _NumbaType(np.int64), _NumbaType(np.float64)
    m.jit = lambda *a, **k: (lambda f: f)
    sys.modules["numba"] = m


def main():
    stub_numba()
    spec = importlib.util.spec_from_file_location("ref_merging_utils", "/root/reference/Utils/merging_utils.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    torch.set_num_threads(1)
    out = {}
    for name, c in cases.merging_cases().items():
        nb, Np, Kl = c["W"].shape
        Ng, Kg = c["S"].shape
        S, W, idx = torch.from_numpy(c["S"]), torch.from_numpy(c["W"]), torch.from_numpy(c["idx"])
        sim = ref.similarity_soft(S, W, idx)                                            # merging_utils.py:6-15
        labels = ref.run_heuristic_solver(sim.numpy(), nb, Kg, Kl)                      # :35-44
        # point2primitive_fusion as evaluation_localSPFN.py:103-110 builds it, then the reference's get_point_final
        A = torch.zeros([Ng, nb * Kl + Kg]).float()
        for b in range(nb):
            A[idx[b], b * Kl:(b + 1) * Kl] = W[b]
        A[:, nb * Kl:] = S
        flag = torch.sum(A[:, :nb * Kl], dim=1) > 0
        A[flag, nb * Kl:] = 0
        fused = ref.get_point_final(A, torch.from_numpy(labels))                        # :46-50
        out[name + "/similarity"] = sim.numpy()
        out[name + "/labels"] = labels.astype(np.int64)
        out[name + "/fused"] = fused.numpy()
    path = os.path.join(ROOT, "tests", "golden", "ref_merging.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
