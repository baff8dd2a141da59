"""Generate tests/golden/ref_fitters.npz by importing the UNMODIFIED reference SPFN
package (CPU, dev container only: needs /root/reference).

    python tests/golden/make_ref_fitters_golden.py

Two import-time shims are applied (no reference file is edited), both required by
torch 2.11:
  * torch.solve was removed      -> torch.linalg.solve (SPFN/geometry_utils.py:140)
  * Tensor.get_device() is -1 on CPU and `.to(-1)` raises -> return the device
    (SPFN/geometry_utils.py:11, SPFN/differentiable_tls.py:10).
Outputs: the parameter dictionary of losses_implementation.compute_parameters for
each case of cases.fitter_cases(), solve_weighted_tls outputs, and the gradients of
sum(x * g) w.r.t. A-weights through Custom_svd_v_colum (TLS backward).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
from tests.golden import cases  # noqa: E402


def install_shims():
    torch.solve = lambda B, A: (torch.linalg.solve(A, B), None)
    orig = torch.Tensor.get_device
    torch.Tensor.get_device = lambda self: self.device if not self.is_cuda else orig(self)


def main():
    install_shims()
    torch.set_num_threads(1)
    from SPFN import differentiable_tls, losses_implementation
    out = {}
    for name, (P, W, X) in cases.fitter_cases().items():
        Pt, Wt, Xt = (torch.from_numpy(a) for a in (P, W, X))
        params = losses_implementation.compute_parameters(Pt, Wt, Xt)
        for k, v in params.items():
            out["%s/%s" % (name, k)] = v.detach().numpy()
    # solve_weighted_tls forward + backward (differentiable_tls.py:200-209, 123-143)
    P, W, X = cases.fitter_cases()["selfcheck"]
    A = torch.from_numpy(P)
    Wt = torch.from_numpy(np.ascontiguousarray(W[:, :, 0])).requires_grad_(True)
    At = A.clone().requires_grad_(True)
    x = differentiable_tls.solve_weighted_tls(At, Wt)
    g = torch.from_numpy(np.random.RandomState(1).randn(*x.shape).astype(np.float32))
    (x * g).sum().backward()
    out["tls/x"] = x.detach().numpy()
    out["tls/g"] = g.numpy()
    out["tls/grad_W"] = Wt.grad.numpy()
    out["tls/grad_A"] = At.grad.numpy()
    # gradients of a scalar loss through all four fitters (reference autograd incl. Custom_svd_v_colum)
    P, W, X = cases.grad_case()
    Pt = torch.from_numpy(P)
    Wt = torch.from_numpy(W).requires_grad_(True)
    Xt = torch.from_numpy(X).requires_grad_(True)
    params = losses_implementation.compute_parameters(Pt, Wt, Xt)
    loss = cases.fitter_loss(params, W, torch)
    loss.backward()
    out["grad/loss"] = np.float64(loss.item())
    out["grad/dW"] = Wt.grad.numpy()
    out["grad/dX"] = Xt.grad.numpy()
    path = os.path.join(ROOT, "tests", "golden", "ref_fitters.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays")


if __name__ == "__main__":
    main()
