"""Generate tests/golden/ref_seg.npz from the UNMODIFIED reference SPFN.losses_implementation (CPU; dev container).

    python tests/golden/make_ref_seg_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
from tests.golden import cases  # noqa: E402


def main():
    from SPFN import losses_implementation as L
    out = {}
    for name, (W, I) in cases.seg_cases().items():
        Wt, It = torch.from_numpy(W), torch.from_numpy(I)
        m = L.hungarian_matching(Wt, It)
        loss, one_minus = L.compute_miou_loss(Wt, It, m)
        out[name + "/matching"] = m.numpy()
        out[name + "/miou_loss"] = loss.numpy()
        out[name + "/one_minus_dot"] = one_minus.numpy()
    path = os.path.join(ROOT, "tests", "golden", "ref_seg.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays")


if __name__ == "__main__":
    main()
