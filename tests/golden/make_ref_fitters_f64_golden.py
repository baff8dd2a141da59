"""Generate tests/golden/ref_fitters_f64.npz: the UNMODIFIED reference fitters and their own autograd (incl.
Custom_svd_v_colum's analytic backward) evaluated in FLOAT64 on cases.grad_case() -- the yardstick the differentiable
path (row a8: cpfn_b200/spfn/_train.py on the moment / small-linalg kernels) is held to at 1e-5.

    python tests/golden/make_ref_fitters_f64_golden.py          (dev container: needs /root/reference)

Besides the two shims of make_ref_fitters_golden.py, the float64 run needs two harness settings (no reference file is
edited): torch's default dtype is float64 while the reference runs, and ``torch.FloatTensor`` -- which
SPFN/geometry_utils.py:16 uses for the three candidate axes -- constructs doubles.  The float32 run of the same code
(tests/golden/ref_fitters.npz, keys grad/*) agrees with this one to 5e-7 (dW) / 1.1e-6 (dX) of the gradient scale.
Outputs: params64/<key>, grad64/loss, grad64/dW, grad64/dX."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
from tests.golden import cases  # noqa: E402
from tests.golden.make_ref_fitters_golden import install_shims  # noqa: E402


def main():
    install_shims()
    torch.set_num_threads(1)
    from SPFN import losses_implementation
    P, W, X = cases.grad_case()
    float_tensor = torch.FloatTensor
    torch.set_default_dtype(torch.float64)
    torch.FloatTensor = torch.DoubleTensor
    try:
        Pt = torch.from_numpy(P).double()
        Wt = torch.from_numpy(W).double().requires_grad_(True)
        Xt = torch.from_numpy(X).double().requires_grad_(True)
        params = losses_implementation.compute_parameters(Pt, Wt, Xt)
        loss = cases.fitter_loss(params, W, torch)
        loss.backward()
    finally:
        torch.FloatTensor = float_tensor
        torch.set_default_dtype(torch.float32)
    out = {"params64/" + k: v.detach().numpy() for k, v in params.items()}
    out["grad64/loss"] = np.float64(loss.item())
    out["grad64/dW"] = Wt.grad.numpy()
    out["grad64/dX"] = Xt.grad.numpy()
    path = os.path.join(ROOT, "tests", "golden", "ref_fitters_f64.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
