"""Device hungarian_matching / compute_miou_loss (csrc/seg_loss.cu, spfn/seg.py) against the reference's outputs
(tests/golden/ref_seg.npz), the numpy + scipy oracle, and -- when the reference Python is staged -- the reference
functions themselves run on the same GPU tensors."""
import os

import numpy as np
import pytest
import torch

from cpfn_b200 import spfn
from oracle import seg as oseg
from tests.golden import cases

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_seg.npz"))


@pytest.mark.parametrize("name", sorted(cases.seg_cases()))
def test_matching_and_miou_against_reference_outputs(cuda_dev, name):
    W, I = cases.seg_cases()[name]
    Wt, It = torch.from_numpy(W).to(cuda_dev), torch.from_numpy(I).to(cuda_dev)
    m = spfn.losses_implementation.hungarian_matching(Wt, It)
    assert m.dtype == torch.int64 and np.array_equal(m.cpu().numpy(), GOLD[name + "/matching"])
    m2, mask = spfn.metric_implementation.hungarian_matching(Wt, It.to(torch.int32))
    want_mask = oseg.hungarian_matching(W, I, with_mask=True)[1]
    assert torch.equal(m2, m) and np.array_equal(mask.cpu().numpy(), want_mask)
    loss, one_minus = spfn.losses_implementation.compute_miou_loss(Wt, It, m)
    assert np.abs(loss.cpu().numpy() - GOLD[name + "/miou_loss"]).max() < 2e-6
    assert np.abs(one_minus.cpu().numpy() - GOLD[name + "/one_minus_dot"]).max() < 2e-6


@pytest.mark.parametrize("seed", range(8))
def test_matching_equals_scipy_on_tie_heavy_costs(cuda_dev, seed):
    """Quantised memberships give cost matrices full of exact ties and zero rows: the assignment must be scipy's."""
    rng = np.random.RandomState(seed)
    B, N, K = 6, 400 + 37 * seed, [8, 12, 21, 28, 32, 5, 16, 24][seed]
    W = (rng.randint(0, 3, size=(B, N, K)) / 2.0).astype(np.float32)
    n_gt = rng.randint(1, K + 1)
    I = rng.randint(-1 if seed % 2 else 0, n_gt, size=(B, N)).astype(np.int64)
    I[:, 0] = n_gt - 1
    got = spfn.seg.hungarian_matching(torch.from_numpy(W).to(cuda_dev), torch.from_numpy(I).to(cuda_dev)).cpu().numpy()
    assert np.array_equal(got, oseg.hungarian_matching(W, I)), seed


def test_miou_gradient(cuda_dev):
    W, I = cases.seg_cases()["background"]
    Wt = torch.from_numpy(W).to(cuda_dev).requires_grad_(True)
    It = torch.from_numpy(I).to(cuda_dev)
    m = spfn.seg.hungarian_matching(Wt, It)
    g = torch.from_numpy(np.random.RandomState(1).randn(*m.shape).astype(np.float32)).to(cuda_dev)
    loss, one_minus = spfn.seg.compute_miou_loss(Wt, It, m)
    ((loss * g).sum() + (one_minus * g).sum()).backward()
    # the reference's formula as torch ops (SPFN/losses_implementation.py:77-89)
    Wr = torch.from_numpy(W).to(cuda_dev).requires_grad_(True)
    B, N, K = W.shape
    n_labels = m.shape[1]
    W_re = torch.gather(Wr, 2, m.unsqueeze(1).expand(B, N, n_labels))
    W_gt = torch.eye(n_labels + 2, device=cuda_dev)[It][:, :, :n_labels]
    dot = torch.sum(W_gt * W_re, axis=1)
    den = torch.sum(W_gt, dim=1) + torch.sum(W_re, dim=1) - dot
    ((1.0 - dot / (den + 1e-10)) * g).sum().add(((1 - dot / N) * g).sum()).backward()
    assert float((Wt.grad - Wr.grad).abs().max()) <= 1e-6 * float(Wr.grad.abs().max()) + 1e-9


def test_against_the_staged_reference_on_the_gpu(cuda_dev):
    from oracle import ref_runtime
    if not ref_runtime.available():
        pytest.skip("baseline/_ref is not staged")
    L = ref_runtime.load_spfn()
    W, I = cases.seg_cases()["soft"]
    Wt, It = torch.from_numpy(W).to(cuda_dev), torch.from_numpy(I).to(cuda_dev)
    want = L.hungarian_matching(Wt, It)
    got = spfn.seg.hungarian_matching(Wt, It)
    assert torch.equal(got, want)
    a, b = L.compute_miou_loss(Wt, It, want)
    c, d = spfn.seg.compute_miou_loss(Wt, It, got)
    assert float((a - c).abs().max()) < 2e-6 and float((b - d).abs().max()) < 2e-6
