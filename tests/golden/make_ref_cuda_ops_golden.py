"""Generate tests/golden/ref_cuda_ops.npz from the UNMODIFIED reference kernels.

Runs on a GPU box only:   python tests/golden/make_ref_cuda_ops_golden.py
Needs oracle/_ref/ref_cuda_ops.so (built in the dev container from /root/reference
by oracle/build_ref.py; the .so travels with the gpurun snapshot).  Writes
gpurun_out/ref_cuda_ops.npz, which is then committed as
tests/golden/ref_cuda_ops.npz.  It stores OUTPUTS of the reference's
farthest_point_sampling / ball_query / three_nn / three_weighted_sum(_grad) /
gather_points(_grad) / group_points(_grad) on the seeded inputs of cases.py.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import build_ref  # noqa: E402
from tests.golden import cases  # noqa: E402


def main():
    ref = build_ref.load_module()
    assert ref is not None, "oracle/_ref/ref_cuda_ops.so missing"
    dev = torch.device("cuda:0")

    def t(a):
        return torch.from_numpy(np.ascontiguousarray(a)).to(dev)

    def fps(xyz, m):
        return ref.farthest_point_sampling(t(xyz), m).cpu().numpy()

    out = {}
    for name, (xyz, m) in cases.fps_cases().items():
        out["fps/" + name] = fps(xyz, m).astype(np.int16 if xyz.shape[1] < 32768 else np.int32)
    for name, (q, xyz, r, k) in cases.ball_cases(fps).items():
        out["ball/" + name] = ref.ball_query(t(q), t(xyz), r, k).cpu().numpy().astype(np.int16)
    for name, (u, kn) in cases.three_nn_cases(fps).items():
        d2, idx = ref.three_nn(t(u), t(kn))
        out["nn_d2/" + name] = d2.cpu().numpy()
        out["nn_idx/" + name] = idx.cpu().numpy().astype(np.int16)
    z = cases.interp_inputs()
    out["tws"] = ref.three_weighted_sum(t(z["pts"]), t(z["idx"]), t(z["w"])).cpu().numpy()
    out["tws_grad"] = ref.three_weighted_sum_grad(t(z["g"]), t(z["idx"]), t(z["w"]), z["M"]).cpu().numpy()
    gi = z["idx"][:, :, 0].copy()
    out["gather"] = ref.gather_points(t(z["pts"]), t(gi)).cpu().numpy()
    out["gather_grad"] = ref.gather_points_grad(t(z["g"]), t(gi), z["M"]).cpu().numpy()
    out["group"] = ref.group_points(t(z["pts"]), t(z["gidx"])).cpu().numpy()
    out["group_grad"] = ref.group_points_grad(t(z["gg"]), t(z["gidx"]), z["M"]).cpu().numpy()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    path = os.path.join(ROOT, "gpurun_out", "ref_cuda_ops.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays;",
          torch.cuda.get_device_name(0))


if __name__ == "__main__":
    main()
