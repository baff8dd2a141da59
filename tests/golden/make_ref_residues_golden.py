"""Generate tests/golden/ref_residues.npz by importing the UNMODIFIED reference SPFN package on CPU (dev container
only: needs /root/reference).

    python tests/golden/make_ref_residues_golden.py
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
    torch.set_num_threads(1)
    from SPFN import losses_implementation, metric_implementation
    params, matching, points, T_gt, P = cases.residue_case()
    pt = {k: torch.from_numpy(v) for k, v in params.items()}
    m, pts, Tg, Pt = (torch.from_numpy(a) for a in (matching, points, T_gt, P))
    out = {}
    loss, per_point = losses_implementation.compute_residue_loss(pt, m, pts, Tg)          # :351-387
    out["residue_loss"], out["residue_per_point"] = loss.numpy(), per_point.numpy()
    loss2, per_point2 = losses_implementation.compute_residue_loss(pt, m, pts, Tg.clamp(max=1), classes=['cone', 'plane'])
    out["residue_loss_cone_plane"], out["residue_per_point_cone_plane"] = loss2.numpy(), per_point2.numpy()
    out["residual"] = metric_implementation.get_residual_loss(pt, m, pts, Tg).numpy()     # :76-81
    for eps in (0.05, 0.2):
        out["p_coverage_%g" % eps] = metric_implementation.compute_P_coverage(Pt, Tg, m, pt, eps).numpy()   # :409-415
    path = os.path.join(ROOT, "tests", "golden", "ref_residues.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()}, {k: out[k] for k in out if k.startswith("p_cov")})


if __name__ == "__main__":
    main()
