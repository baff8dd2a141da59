"""CPU check of the differentiable fitter algebra (cpfn_b200/spfn/_train.py).

The product computes the weighted moments with a CUDA kernel; here, and ONLY here, that one
linear map is replaced by a float64 torch einsum so that the moment identities, the custom SVD
backward and the guarded solves can be checked without a GPU:
  * forward against the numpy oracle (oracle/fitters.py, pinned to the reference SPFN package);
  * gradients of a scalar loss w.r.t. W and X against the reference's own autograd
    (tests/golden/ref_fitters.npz, keys grad/*, made by make_ref_fitters_golden.py).
The GPU test (tests/test_gpu_fitters.py) repeats both with the real kernels."""
import os

import numpy as np
import pytest
import torch

from cpfn_b200.spfn import _train
from oracle import fitters as ofit
from tests.golden import cases

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ref_fitters.npz")


def torch_moments(Wt, P, X):
    """Test-only restatement of cpfn_weighted_moments (same feature order), float64, differentiable."""
    Wt, P, X = Wt.double(), P.double(), X.double()
    iu = [(0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2)]
    it = [(0, 0, 0), (0, 0, 1), (0, 0, 2), (0, 1, 1), (0, 1, 2), (0, 2, 2), (1, 1, 1), (1, 1, 2), (1, 2, 2), (2, 2, 2)]
    px = (P * X).sum(-1, keepdim=True)
    feats = [torch.ones_like(P[..., :1]), P] + [P[..., i:i + 1] * P[..., j:j + 1] for i, j in iu] \
        + [P[..., i:i + 1] * P[..., j:j + 1] * P[..., k:k + 1] for i, j, k in it] + [X] \
        + [X[..., i:i + 1] * X[..., j:j + 1] for i, j in iu] + [X * px]
    psi = torch.cat(feats, dim=-1)                       # [B,N,32]
    return torch.einsum("bnk,bnf->bkf", Wt, psi)


def torch_sym_eigh(M, vectors=True):
    """Test-only restatement of cpfn_sym_eigh_small (the contract of torch.linalg.eigh)."""
    lam, Q = torch.linalg.eigh(M.detach().double())
    return lam, (Q if vectors else None)


def torch_small_solve(A, b):
    """Test-only restatement of cpfn_small_solve (differentiable through torch)."""
    return torch.linalg.solve(A, b.unsqueeze(-1)).squeeze(-1)


@pytest.fixture()
def cpu_moments(monkeypatch):
    monkeypatch.setattr(_train, "weighted_moments", torch_moments)
    monkeypatch.setattr(_train, "sym_eigh", torch_sym_eigh)
    monkeypatch.setattr(_train, "small_solve", torch_small_solve)


def _signfix(a, b):
    return a * np.sign(np.sum(a * b, axis=-1, keepdims=True))


def test_forward_matches_oracle(cpu_moments):
    P, W, X = cases.fitter_cases()["shape_2048_k24"]
    got = _train.compute_parameters(torch.from_numpy(P), torch.from_numpy(W), torch.from_numpy(X),
                                    ("plane", "sphere", "cylinder", "cone"))
    ref = ofit.compute_parameters(P, W, X)
    for key, val in got.items():
        a, b = val.detach().numpy().astype(np.float64), ref[key].astype(np.float64)
        m = cases.fit_mask("shape_2048_k24", key, W)
        if key in ("plane_normal", "cylinder_axis"):
            a = _signfix(a, b)
        if key == "plane_center":
            a = a * np.sign(np.sum(got["plane_normal"].detach().numpy() * ref["plane_normal"], axis=-1))
        assert np.abs(a[m] - b[m]).max() <= 1e-5 * max(1.0, np.abs(b[m]).max()), key


def test_gradients_match_reference_autograd(cpu_moments):
    g = np.load(GOLDEN)
    P, W, X = cases.grad_case()
    Wt = torch.from_numpy(W).requires_grad_(True)
    Xt = torch.from_numpy(X).requires_grad_(True)
    params = _train.compute_parameters(torch.from_numpy(P), Wt, Xt, ("plane", "sphere", "cylinder", "cone"))
    loss = cases.fitter_loss(params, W, torch)
    loss.backward()
    assert abs(loss.item() - float(g["grad/loss"])) <= 1e-4 * max(1.0, abs(float(g["grad/loss"])))
    for name, got in (("dW", Wt.grad.numpy()), ("dX", Xt.grad.numpy())):
        ref = g["grad/" + name]
        err = np.abs(got - ref).max() / np.abs(ref).max()
        assert err < 1e-5, (name, err)          # fp32 reference run (5e-7 / 1.1e-6 from its own fp64 run)
    g64 = np.load(os.path.join(os.path.dirname(GOLDEN), "ref_fitters_f64.npz"))       # the reference in float64
    assert abs(loss.item() - float(g64["grad64/loss"])) <= 1e-6
    for name, got in (("dW", Wt.grad.numpy()), ("dX", Xt.grad.numpy())):
        ref = g64["grad64/" + name]
        assert np.abs(got - ref).max() / np.abs(ref).max() < 2e-6, name
    for key, val in params.items():
        a, b = val.detach().numpy().astype(np.float64), g64["params64/" + key]
        if key in ("plane_normal", "cylinder_axis"):
            a = _signfix(a, b)
        if key == "plane_center":
            a = a * np.sign(np.sum(params["plane_normal"].detach().numpy() * g64["params64/plane_normal"], axis=-1))
        assert np.abs(a - b).max() <= 2e-6 * max(1.0, np.abs(b).max()), key


def test_svd_column_backward_formula(cpu_moments):
    """Custom_svd_v_colum backward (differentiable_tls.py:132-143) vs finite differences on a well
    separated symmetric matrix (the reference's own gradcheck, :162-176, in float64)."""
    torch.manual_seed(0)
    A = torch.randn(4, 3, 3, dtype=torch.float64)
    M = (A @ A.transpose(1, 2) + torch.diag_embed(torch.tensor([3.0, 1.0, 0.2], dtype=torch.float64))).requires_grad_(True)
    gvec = torch.randn(4, 3, dtype=torch.float64)
    f = lambda m: (_train.svd_v_last_column(0.5 * (m + m.transpose(1, 2))) * gvec).sum()
    f(M).backward()
    num = torch.zeros_like(M)
    eps = 1e-6
    with torch.no_grad():
        for idx in np.ndindex(*M.shape):
            d = torch.zeros_like(M); d[idx] = eps
            num[idx] = (f(M + d) - f(M - d)) / (2 * eps)
    assert (M.grad - num).abs().max() < 1e-6
