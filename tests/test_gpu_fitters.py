"""GPU parity of the fused primitive fitters (cpfn_fit_primitives through the SPFN-named
Python interface) against the numpy oracle (oracle/fitters.py, pinned to the reference
SPFN package) and against the committed reference outputs (tests/golden/ref_fitters.npz).

Tolerance (north_star: 1e-5 for fp32 TLS): 1e-5 relative to the output scale
(max(1, |ref|_inf)) on the slots where the fit is well posed (cases.fit_mask); the
pure-noise `selfcheck` case is ill-conditioned by construction and gets 2e-3, the same
margin the oracle itself needs against the reference there."""
import os

import numpy as np
import pytest
import torch

from cpfn_b200 import spfn
from oracle import fitters as ofit
from tests.golden import cases

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ref_fitters.npz")
SIGN_FREE = {"plane_normal": None, "cylinder_axis": None}
FLIPS_WITH = {"plane_center": "plane_normal"}


def _run(P, W, X, dev):
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    with torch.no_grad():
        out = spfn.losses_implementation.compute_parameters(t(P), t(W), t(X))
    return {k: v.cpu().numpy() for k, v in out.items()}


def _compare(got, ref, name, W, tol):
    worst = {}
    for key, val in got.items():
        live = cases.fit_mask(name, key, W)
        a, b = val.astype(np.float64).copy(), ref[key].astype(np.float64)
        if key in SIGN_FREE:
            a = a * np.sign(np.sum(a * b, axis=-1, keepdims=True))
        if key in FLIPS_WITH:
            o = FLIPS_WITH[key]
            a = a * np.sign(np.sum(got[o].astype(np.float64) * ref[o], axis=-1))
        m = live if a.ndim == 2 else np.broadcast_to(live[..., None], a.shape)
        if not m.any():
            continue
        scale = max(1.0, float(np.abs(b[m]).max()))
        worst[key] = float(np.abs(a[m] - b[m]).max() / scale)
    bad = {k: v for k, v in worst.items() if not v <= tol}
    assert not bad, (name, bad)
    return worst


CASES = ["selfcheck", "shape_2048_k24", "config1_8192_k24", "onehot_4096_k28"]


@pytest.mark.parametrize("name", CASES)
def test_fit_matches_oracle(cuda_dev, name):
    P, W, X = cases.fitter_cases()[name]
    got = _run(P, W, X, cuda_dev)
    ref = ofit.compute_parameters(P, W, X)
    _compare(got, ref, name, W, 2e-3 if name == "selfcheck" else 1e-5)


@pytest.mark.parametrize("name", CASES)
def test_fit_matches_reference_golden(cuda_dev, name):
    g = np.load(GOLDEN)
    P, W, X = cases.fitter_cases()[name]
    got = _run(P, W, X, cuda_dev)
    ref = {k: g["%s/%s" % (name, k)] for k in got}
    _compare(got, ref, name, W, 2e-3 if name == "selfcheck" else 2e-5)


def test_individual_fitters_and_keys(cuda_dev):
    P, W, X = cases.fitter_cases()["shape_2048_k24"]
    t = lambda a: torch.from_numpy(a).to(cuda_dev)
    with torch.no_grad():
        n, c = spfn.plane_fitter.compute_parameters(t(P), t(W))
        sc, sr = spfn.sphere_fitter.compute_parameters(t(P), t(W))
        ca, cc, cr = spfn.cylinder_fitter.compute_parameters(t(P), t(W), t(X))
        ap, ax, ha = spfn.cone_fitter.compute_parameters(t(P), t(W), t(X))
        full = spfn.losses_implementation.compute_parameters(t(P), t(W), t(X))
        two = spfn.losses_implementation.compute_parameters(t(P), t(W), t(X), classes=["sphere", "cone"])
    B, _, K = W.shape
    assert n.shape == (B, K, 3) and c.shape == (B, K) and ha.shape == (B, K)
    assert n.is_contiguous() and cc.is_contiguous()
    assert list(two.keys()) == ["sphere_center", "sphere_radius_squared", "cone_apex", "cone_axis", "cone_half_angle"]
    for a, b in ((n, full["plane_normal"]), (sc, full["sphere_center"]), (ca, full["cylinder_axis"]),
                 (cc, full["cylinder_center"]), (ap, full["cone_apex"]), (ha, full["cone_half_angle"])):
        assert torch.equal(a, b)          # deterministic: no atomics anywhere in the fitters
    with pytest.raises(NotImplementedError):
        spfn.losses_implementation.compute_parameters(t(P), t(W), t(X), classes=["torus"])
    with pytest.raises(RuntimeError):
        spfn.plane_fitter.compute_parameters(torch.from_numpy(P), torch.from_numpy(W))


def test_fit_properties_full_size(cuda_dev):
    """BASELINE config sizes (B=16, N=8192, K=28): size-independent properties -- unit normals /
    axes, exact recovery of an analytic sphere and plane, invariance to point order."""
    from cpfn_b200 import synth
    P, X, W, _ = synth.shape_batch(16, 8192, seed=1235, k_slots=28)
    got = _run(P, W, X, cuda_dev)
    for k in ("plane_normal", "cylinder_axis", "cone_axis"):
        np.testing.assert_allclose(np.linalg.norm(got[k], axis=-1), 1.0, atol=1e-5)
    perm = np.random.default_rng(0).permutation(8192)
    got2 = _run(P[:, perm], W[:, perm], X[:, perm], cuda_dev)
    for k in got:
        m = cases.fit_mask("shape", k, W)
        a, b = got[k][m], got2[k][m]
        assert np.abs(np.abs(a) - np.abs(b)).max() < 2e-5, k
    # analytic sphere: every point on |p - c0| = 0.5, one slot with all the weight
    rng = np.random.default_rng(1)
    d = rng.normal(size=(1, 8192, 3)); d /= np.linalg.norm(d, axis=2, keepdims=True)
    c0 = np.array([0.1, -0.2, 0.3])
    Ps = (c0 + 0.5 * d).astype(np.float32)
    Ws = np.zeros((1, 8192, 4), np.float32); Ws[:, :, 1] = 1.0
    gs = _run(Ps, Ws, d.astype(np.float32), cuda_dev)
    np.testing.assert_allclose(gs["sphere_center"][0, 1], c0, atol=2e-6)
    np.testing.assert_allclose(gs["sphere_radius_squared"][0, 1], 0.25, atol=2e-6)
    ref = ofit.compute_parameters(Ps, Ws, d.astype(np.float32))
    np.testing.assert_allclose(gs["cone_apex"][0, 1], ref["cone_apex"][0, 1], atol=1e-5)


def test_fit_edge_cases(cuda_dev):
    """Ragged / tiny sizes and empty slots must not crash or produce NaN in live slots."""
    rng = np.random.default_rng(3)
    for (B, N, K) in [(1, 1, 1), (1, 7, 3), (3, 33, 5), (2, 1000, 21), (1, 5000, 40), (1, 300, 256)]:
        P = rng.normal(size=(B, N, 3)).astype(np.float32)
        X = rng.normal(size=(B, N, 3)); X = (X / np.linalg.norm(X, axis=2, keepdims=True)).astype(np.float32)
        W = rng.uniform(size=(B, N, K)).astype(np.float32)
        got = _run(P, W, X, cuda_dev)
        if N >= 33 and K <= 40:
            ref = ofit.compute_parameters(P, W, X)
            for k in ("plane_center", "sphere_radius_squared", "cone_half_angle"):
                a, b = np.abs(got[k]), np.abs(ref[k])
                assert np.abs(a - b).max() <= 5e-3 * max(1.0, b.max()), (B, N, K, k)
        assert all(np.isfinite(v).all() for v in got.values()) or N < 4
    assert spfn.fit.fit_primitives(torch.zeros(0, 16, 3, device=cuda_dev), torch.zeros(0, 16, 4, device=cuda_dev),
                                   torch.zeros(0, 16, 3, device=cuda_dev))["plane_center"].shape == (0, 4)


# ---- training path (CUDA moment kernels + float64 autograd algebra) ---------------------------

def test_weighted_moments_forward_and_backward(cuda_dev):
    from cpfn_b200.spfn import _train
    from tests.test_train_algebra_cpu import torch_moments
    rng = np.random.default_rng(4)
    for (B, N, K) in [(2, 1000, 12), (1, 333, 1), (3, 64, 40)]:
        P = torch.from_numpy(rng.normal(size=(B, N, 3)).astype(np.float32)).to(cuda_dev)
        X = torch.from_numpy(rng.normal(size=(B, N, 3)).astype(np.float32)).to(cuda_dev).requires_grad_(True)
        W = torch.from_numpy(rng.uniform(size=(B, N, K)).astype(np.float32)).to(cuda_dev).requires_grad_(True)
        G = torch.from_numpy(rng.normal(size=(B, K, 32))).to(cuda_dev)
        M = _train.weighted_moments(W, P, X)
        (M * G).sum().backward()
        dW, dX = W.grad.clone(), X.grad.clone()
        W.grad = None; X.grad = None
        Mr = torch_moments(W, P, X)
        (Mr * G).sum().backward()
        assert float((M - Mr).abs().max() / Mr.abs().max()) < 1e-6
        assert float((dW - W.grad).abs().max() / W.grad.abs().max()) < 1e-5
        assert float((dX - X.grad).abs().max() / X.grad.abs().max()) < 1e-5


def test_training_path_forward_equals_inference_and_reference_gradients(cuda_dev):
    g = np.load(GOLDEN)
    P, W, X = cases.grad_case()
    t = lambda a: torch.from_numpy(a).to(cuda_dev)
    with torch.no_grad():
        inf = spfn.losses_implementation.compute_parameters(t(P), t(W), t(X))
    Wt, Xt = t(W).requires_grad_(True), t(X).requires_grad_(True)
    tr = spfn.losses_implementation.compute_parameters(t(P), Wt, Xt)
    ref = {k: v.cpu().numpy() for k, v in inf.items()}
    _compare({k: v.detach().cpu().numpy() for k, v in tr.items()}, ref, "grad", W, 1e-5)
    loss = cases.fitter_loss(tr, W, torch)
    loss.backward()
    assert abs(loss.item() - float(g["grad/loss"])) <= 1e-4 * max(1.0, abs(float(g["grad/loss"])))
    for name, got in (("dW", Wt.grad.cpu().numpy()), ("dX", Xt.grad.cpu().numpy())):
        err = np.abs(got - g["grad/" + name]).max() / np.abs(g["grad/" + name]).max()
        assert err < 1e-5, (name, err)       # the reference's fp32 run is 5e-7 / 1.1e-6 from its own fp64 run


def test_small_linalg_kernels(cuda_dev):
    """cpfn_sym_eigh_small / cpfn_small_solve (csrc/small_linalg.cu) against torch.linalg in float64: well separated,
    nearly degenerate, rank-deficient and indefinite symmetric matrices; solves incl. the transposed system and the
    autograd wrapper's backward."""
    from cpfn_b200.spfn import _train
    torch.manual_seed(3)
    for D in (2, 3):
        A = torch.randn(500, D, D, dtype=torch.float64, device=cuda_dev)
        S = A @ A.transpose(1, 2)
        S[100:200] *= torch.logspace(-12, 3, 100, dtype=torch.float64, device=cuda_dev)[:, None, None]
        S[200:300] = A[200:300] + A[200:300].transpose(1, 2)                       # indefinite
        v = torch.randn(100, D, 1, dtype=torch.float64, device=cuda_dev)
        S[300:400] = v @ v.transpose(1, 2)                                         # rank one
        S[400:450] = torch.eye(D, dtype=torch.float64, device=cuda_dev) * 2.5      # fully degenerate
        lam, Q = _train.sym_eigh(S)
        lam_ref = torch.linalg.eigvalsh(S)
        scale = lam_ref.abs().max(dim=1, keepdim=True)[0].clamp(min=1e-300)
        assert float(((lam - lam_ref).abs() / scale).max()) < 1e-13
        assert float((Q.transpose(1, 2) @ Q - torch.eye(D, dtype=torch.float64, device=cuda_dev)).abs().max()) < 1e-13
        assert float(((S @ Q - Q * lam.unsqueeze(1)).abs() / scale.unsqueeze(1)).max()) < 1e-12
        assert torch.equal(_train.sym_eigh(S, vectors=False)[0], lam)
        Bm = (S[:100] + torch.eye(D, dtype=torch.float64, device=cuda_dev)).clone().requires_grad_(True)
        b = torch.randn(100, D, dtype=torch.float64, device=cuda_dev, requires_grad=True)
        g = torch.randn(100, D, dtype=torch.float64, device=cuda_dev)
        x = _train.small_solve(Bm, b)
        (x * g).sum().backward()
        dB, db = Bm.grad.clone(), b.grad.clone()
        Bm.grad = None; b.grad = None
        xr = torch.linalg.solve(Bm, b.unsqueeze(-1)).squeeze(-1)
        (xr * g).sum().backward()
        assert float((x - xr).abs().max() / xr.abs().max()) < 1e-12
        assert float((dB - Bm.grad).abs().max() / Bm.grad.abs().max()) < 1e-11
        assert float((db - b.grad).abs().max() / b.grad.abs().max()) < 1e-11
        An = torch.randn(64, D, D, dtype=torch.float64, device=cuda_dev)              # non-symmetric, needs pivoting
        An[:, 0, 0] = 1e-14
        bn = torch.randn(64, D, dtype=torch.float64, device=cuda_dev)
        xs = _train._solve_raw(An, bn, True)
        assert float((An.transpose(1, 2) @ xs.unsqueeze(-1) - bn.unsqueeze(-1)).abs().max()) < 1e-9
    with pytest.raises(RuntimeError):
        _train.sym_eigh(torch.eye(3, dtype=torch.float64).unsqueeze(0))              # no CPU path


def test_training_path_against_float64_reference(cuda_dev):
    """Row a8 at the north star's 1e-5: the differentiable fitters (moment kernels + small-linalg kernels + float64
    identities) against the UNMODIFIED reference evaluated in float64 (tests/golden/ref_fitters_f64.npz,
    make_ref_fitters_f64_golden.py): all ten parameter tensors, the loss, and the gradients w.r.t. W and X that the
    reference's own autograd -- Custom_svd_v_colum's analytic backward included -- produces."""
    g = np.load(os.path.join(os.path.dirname(GOLDEN), "ref_fitters_f64.npz"))
    P, W, X = cases.grad_case()
    t = lambda a: torch.from_numpy(a).to(cuda_dev)
    Wt, Xt = t(W).requires_grad_(True), t(X).requires_grad_(True)
    tr = spfn.losses_implementation.compute_parameters(t(P), Wt, Xt)
    for key, val in tr.items():
        a, b = val.detach().cpu().numpy().astype(np.float64), g["params64/" + key]
        if key in ("plane_normal", "cylinder_axis"):
            a = a * np.sign(np.sum(a * b, axis=-1, keepdims=True))
        if key == "plane_center":
            a = a * np.sign(np.sum(tr["plane_normal"].detach().cpu().numpy() * g["params64/plane_normal"], axis=-1))
        m = cases.fit_mask("grad", key, W)            # well-posed slots (the ones the loss and its gradient see)
        assert np.abs(a[m] - b[m]).max() <= 1e-5 * max(1.0, np.abs(b[m]).max()), (key, np.abs(a[m] - b[m]).max())
    loss = cases.fitter_loss(tr, W, torch)
    loss.backward()
    assert abs(loss.item() - float(g["grad64/loss"])) <= 1e-5 * max(1.0, abs(float(g["grad64/loss"])))
    for name, got in (("dW", Wt.grad.cpu().numpy()), ("dX", Xt.grad.cpu().numpy())):
        ref = g["grad64/" + name]
        err = np.abs(got - ref).max() / np.abs(ref).max()
        assert err < 1e-5, (name, err)


def test_b3_function_api(cuda_dev):
    """solve_weighted_tls / weighted_plane_fitting / weighted_sphere_fitting / guarded_matrix_solve_ls /
    compute_consistent_plane_frame against the numpy oracle on the reference's self-check inputs."""
    P, W, X = cases.fitter_cases()["selfcheck"]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda_dev)
    w0 = np.ascontiguousarray(W[:, :, 0])
    x = spfn.differentiable_tls.solve_weighted_tls(t(P), t(w0)).cpu().numpy()
    xr = ofit.solve_weighted_tls(P, w0)
    np.testing.assert_allclose(x * np.sign(np.sum(x * xr, 1, keepdims=True)), xr, atol=2e-3)
    g = np.load(GOLDEN)
    A = t(P).requires_grad_(True)
    Wt = t(w0).requires_grad_(True)
    xx = spfn.differentiable_tls.solve_weighted_tls(A, Wt)
    s = torch.sign((xx.detach() * t(g["tls/x"])).sum(1, keepdim=True))
    (xx * s * t(g["tls/g"])).sum().backward()
    for got, ref in ((Wt.grad, g["tls/grad_W"]), (A.grad, g["tls/grad_A"])):
        err = np.abs(got.cpu().numpy() - ref).max() / np.abs(ref).max()
        assert err < 5e-3, err
    n, c = spfn.geometry_utils.weighted_plane_fitting(t(P), t(w0))
    nr, cr = ofit.weighted_plane_fitting(P, w0)
    sg = np.sign(np.sum(n.cpu().numpy() * nr, 1))
    np.testing.assert_allclose(n.cpu().numpy() * sg[:, None], nr, atol=2e-3)
    np.testing.assert_allclose(c.cpu().numpy() * sg, cr, atol=2e-3)
    ce, r2 = spfn.geometry_utils.weighted_sphere_fitting(t(P), t(w0))
    cer, r2r = ofit.weighted_sphere_fitting(P, w0)
    np.testing.assert_allclose(ce.cpu().numpy(), cer, atol=2e-3)
    np.testing.assert_allclose(r2.cpu().numpy(), r2r, rtol=2e-3)
    ce2, r22 = spfn.geometry_utils.weighted_sphere_fitting(t(P[:, :, :2]), t(w0))
    ce2r, r22r = ofit.weighted_sphere_fitting(P[:, :, :2], w0)
    np.testing.assert_allclose(ce2.cpu().numpy(), ce2r, atol=2e-3)
    b = (P ** 2).sum(2, keepdims=True).astype(np.float32)
    xs = spfn.geometry_utils.guarded_matrix_solve_ls(t(P), t(b), t(w0)).cpu().numpy()
    np.testing.assert_allclose(xs, ofit.guarded_matrix_solve_ls(P, b, w0), atol=2e-3)
    xa, ya = spfn.geometry_utils.compute_consistent_plane_frame(t(nr))
    xar, yar = ofit.compute_consistent_plane_frame(nr)
    np.testing.assert_allclose(xa.cpu().numpy(), xar, atol=1e-6)
    np.testing.assert_allclose(ya.cpu().numpy(), yar, atol=1e-6)
