"""GPU parity of the patch -> object merge through the C ABI: similarity vs the exactly
rounded (fp64) oracle and the goldens of the unmodified reference, merged labels, fused memberships, merged
normals / types (bit-exact: same summation order as scatter_add_ on the CPU), and the same checks against a
dense fp64 torch restatement at the full LocalSPFN size (131 072 points, 32 patches of 8192, 21 + 28 slots)."""
import os

import numpy as np
import pytest
import torch

from cpfn_b200 import merging_utils
from oracle import merging as omerge
from tests.golden import cases

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_merging.npz"))
CASES = cases.merging_cases()


def _dev(c, dev):
    return {k: torch.from_numpy(v).to(dev) for k, v in c.items()}


@pytest.mark.parametrize("name", sorted(CASES))
def test_similarity_labels_fusion(name, cuda_dev):
    c, d = CASES[name], _dev(CASES[name], cuda_dev)
    nb, Np, Kl = c["W"].shape
    Kg = c["S"].shape[1]
    sim = merging_utils.similarity_soft(d["S"], d["W"], d["idx"]).cpu().numpy()
    exact = omerge.similarity_soft(c["S"], c["W"], c["idx"], dtype=np.float64)
    scale = np.abs(exact).max()
    assert sim.dtype == np.float32 and np.array_equal(sim, sim.T)
    assert np.abs(sim - exact).max() <= 1e-6 * scale and np.allclose(sim, exact, rtol=1e-6, atol=1e-6)
    assert np.abs(sim - GOLD[name + "/similarity"]).max() <= 1e-5 * scale      # the reference's fp32 mm
    labels = merging_utils.run_heuristic_solver(sim, nb, Kg, Kl)
    assert np.array_equal(labels, GOLD[name + "/labels"])
    fused = merging_utils.fuse_patches(d["S"], d["W"], d["idx"], labels).cpu().numpy()
    assert fused.shape == GOLD[name + "/fused"].shape and np.abs(fused - GOLD[name + "/fused"]).max() <= 1e-6
    # the reference's own call: dense operand
    A = omerge.point2primitive(c["S"], c["W"], c["idx"], False)
    A[np.sum(A[:, :nb * Kl], axis=1) > 0, nb * Kl:] = 0
    dense = merging_utils.get_point_final(torch.from_numpy(A).to(cuda_dev), torch.from_numpy(labels)).cpu().numpy()
    assert np.abs(dense - GOLD[name + "/fused"]).max() <= 1e-6
    assert np.array_equal(dense, fused)                      # same order of additions in both kernels


@pytest.mark.parametrize("name", sorted(CASES))
def test_normals_and_types_bit_exact(name, cuda_dev):
    c, d = CASES[name], _dev(CASES[name], cuda_dev)
    Xg, Tg = merging_utils.merge_normals_types(d["X"], d["T"], d["idx"], d["obj_normals"], d["obj_types"])
    Xo, To = omerge.merge_normals_types(c["X"], c["T"], c["idx"], c["obj_normals"], c["obj_types"])
    assert np.array_equal(Xg.cpu().numpy(), Xo) and np.array_equal(Tg.cpu().numpy(), To)


def test_merge_shape_pipeline_and_errors(cuda_dev):
    c, d = CASES["wide"], _dev(CASES["wide"], cuda_dev)
    Wf, Xg, Tg, labels = merging_utils.merge_shape(d["W"], d["X"], d["T"], d["idx"], d["S"], d["obj_normals"], d["obj_types"])
    assert np.array_equal(labels, GOLD["wide/labels"]) and np.abs(Wf.cpu().numpy() - GOLD["wide/fused"]).max() <= 1e-6
    assert Xg.shape == (6000, 3) and Tg.shape == (6000, 4)
    with pytest.raises(RuntimeError):
        merging_utils.similarity_soft(d["S"].cpu(), d["W"].cpu(), d["idx"].cpu())
    with pytest.raises(RuntimeError):                        # more than 32 slots per patch is not supported
        merging_utils.similarity_soft(d["S"], torch.rand(2, 64, 40, device=cuda_dev), torch.arange(128, device=cuda_dev).view(2, 64))


def test_full_size_against_dense_fp64(cuda_dev):
    Ng, nb, Np, Kl, Kg = 131072, 32, 8192, 21, 28
    g = torch.Generator(device="cpu").manual_seed(5)
    P = torch.from_numpy(cases.synth.shape_cloud(Ng, 99)[0]).to(cuda_dev)
    idx = torch.empty(nb, Np, dtype=torch.int64, device=cuda_dev)
    for b in range(nb):
        c = P[int(torch.randint(Ng, (1,), generator=g))]
        idx[b] = torch.topk((P - c).norm(dim=1), Np, largest=False).indices
    W = torch.softmax(4 * torch.randn(nb, Np, Kl, generator=g).to(cuda_dev), dim=2)
    S = torch.nn.functional.one_hot(torch.randint(Kg, (Ng,), generator=g), Kg).to(cuda_dev)
    X = torch.nn.functional.normalize(torch.randn(nb, Np, 3, generator=g), dim=2).to(cuda_dev)
    T = torch.randn(nb, Np, 4, generator=g).to(cuda_dev)
    on = torch.nn.functional.normalize(torch.randn(Ng, 3, generator=g), dim=1).to(cuda_dev)
    ot = torch.randn(Ng, 4, generator=g).to(cuda_dev)
    inverse = merging_utils.inverse_index(idx, Ng)
    sim = merging_utils.similarity_soft(S, W, idx, inverse=inverse)
    M = nb * Kl + Kg
    A = torch.zeros(Ng, M, dtype=torch.float64, device=cuda_dev)
    for b in range(nb):
        A[idx[b], b * Kl:(b + 1) * Kl] += W[b].double()
    A[:, nb * Kl:] = S.double()
    exact = A.T @ A
    assert float((sim.double() - exact).abs().max()) <= 1e-6 * float(exact.max())
    assert torch.equal(sim, sim.T)
    labels = merging_utils.run_heuristic_solver(sim.cpu().numpy(), nb, Kg, Kl)
    fused = merging_utils.fuse_patches(S, W, idx, labels, inverse=inverse)
    covered = A[:, :nb * Kl].sum(1) > 0
    A[covered, nb * Kl:] = 0
    lab = torch.from_numpy(labels).to(cuda_dev)
    onehot = torch.nn.functional.one_hot(lab).double()
    want = A @ (onehot / (onehot.sum(0, keepdim=True) + 1e-10))
    assert fused.shape == want.shape and float((fused.double() - want).abs().max()) <= 2e-6
    assert 0.3 < float(covered.float().mean()) < 1.0
    Xg, Tg = merging_utils.merge_normals_types(X, T, idx, on, ot, inverse=inverse)
    flat = idx.view(-1)
    Xs = torch.zeros(Ng, 3, dtype=torch.float64, device=cuda_dev).index_add_(0, flat, X.view(-1, 3).double())
    Xs[~covered] = on[~covered].double()
    assert float((Xg.double() - torch.nn.functional.normalize(Xs, dim=1)).abs().max()) <= 1e-5
    num = torch.zeros(Ng, 4, dtype=torch.float64, device=cuda_dev).index_add_(0, flat, T.view(-1, 4).double())
    den = torch.zeros(Ng, dtype=torch.float64, device=cuda_dev).index_add_(0, flat, torch.ones_like(flat, dtype=torch.float64))
    Tw = num / den.clamp(min=1)[:, None]
    Tw[~covered] = ot[~covered].double()
    assert float((Tg.double() - Tw).abs().max()) <= 1e-5


def _device_solve(sim_np, nb, Kg, Kl, dev, threshold=0):
    sim = torch.from_numpy(np.ascontiguousarray(sim_np, dtype=np.float32)).to(dev)
    labels, weights, n_labels, seg = merging_utils.solve_labels_device(sim, nb, Kg, Kl, threshold=threshold,
                                                                       return_segments=True)
    L = int(n_labels.item())
    return labels.cpu().numpy().astype(np.int64), weights.cpu().numpy(), L, seg.cpu().numpy().astype(np.int64)


@pytest.mark.parametrize("name", sorted(CASES))
def test_device_solver_matches_reference_labels(name, cuda_dev):
    """The on-device label solve (csrc/merge_solve.cu) against the labels of the UNMODIFIED reference
    (tests/golden/ref_merging.npz) and the host C pass, from the same float32 similarity matrix."""
    c, d = CASES[name], _dev(CASES[name], cuda_dev)
    nb, Np, Kl = c["W"].shape
    Kg = c["S"].shape[1]
    sim = merging_utils.similarity_soft(d["S"], d["W"], d["idx"])
    labels, weights, L, _ = _device_solve(sim.cpu().numpy(), nb, Kg, Kl, cuda_dev)
    assert np.array_equal(labels, GOLD[name + "/labels"])
    host = merging_utils.run_heuristic_solver(sim.cpu().numpy(), nb, Kg, Kl)
    assert np.array_equal(labels, host) and L == host.max() + 1
    counts = np.bincount(host, minlength=L).astype(np.float32)
    assert np.array_equal(weights[:L], np.float32(1) / (counts + np.float32(1e-10))) and not weights[L:].any()
    Wf, Xg, Tg, lab = merging_utils.merge_shape(d["W"], d["X"], d["T"], d["idx"], d["S"], d["obj_normals"], d["obj_types"])
    Wh, Xh, Th, labh = merging_utils.merge_shape(d["W"], d["X"], d["T"], d["idx"], d["S"], d["obj_normals"],
                                                 d["obj_types"], solver="host")
    assert np.array_equal(lab, labh) and torch.equal(Wf, Wh) and torch.equal(Xg, Xh) and torch.equal(Tg, Th)


@pytest.mark.parametrize("seed", range(12))
def test_device_solver_random_matrices(seed, cuda_dev):
    """Random symmetric matrices with the structures that steer the greedy pass: many exact ties (quantised
    values), a first arg-max inside one patch, empty slots (diagonal below the threshold), no pair at all, a
    threshold above most entries -- against the literal repeated-argmax loop of the reference (oracle) and the
    host pass; raw segment ids included."""
    rng = np.random.RandomState(100 + seed)
    nb, Kl, Kg = [(3, 4, 5), (7, 6, 7), (32, 21, 28), (1, 5, 4), (12, 3, 9), (63, 2, 3)][seed % 6]
    M = nb * Kl + Kg
    A = rng.rand(M, M).astype(np.float32)
    if seed % 3 == 0:
        A = np.round(A * 8) / 8                                   # ties everywhere
    sim = ((A + A.T) * (rng.rand(M, M) < (0.3 if seed % 2 else 0.9))).astype(np.float32)
    sim = np.maximum(sim, sim.T)
    np.fill_diagonal(sim, rng.rand(M).astype(np.float32) * (rng.rand(M) < 0.8))     # some empty slots (diagonal 0)
    thr = [0, 0, 0.5, 0, 1.5, 0][seed % 6]
    if seed == 4:
        sim[1, 2] = sim[2, 1] = 50.0                              # the first arg-max lies inside patch 0
    if seed == 5:
        sim[~np.eye(M, dtype=bool)] = 0.0                         # no pair above the threshold
    labels, weights, L, seg = _device_solve(sim, nb, Kg, Kl, cuda_dev, threshold=thr)
    host = merging_utils.run_heuristic_solver(sim, nb, Kg, Kl, threshold=thr)
    assert np.array_equal(labels, host), (seed, np.flatnonzero(labels != host)[:10])
    if M <= 200:                                                  # the literal loop is O(pairs x merges)
        want = omerge.run_heuristic_solver(sim, nb, Kg, Kl, threshold=thr)
        assert np.array_equal(labels, want)
    idx = np.where(sim > np.float32(thr))
    keep = idx[0] < idx[1]
    pairs = np.stack([idx[0][keep], idx[1][keep]], axis=1).astype(np.int64)
    patch_id = np.concatenate((np.repeat(np.arange(nb), Kl), nb * np.ones(Kg, dtype=int))).astype(np.int64)
    if len(pairs):
        seg_host = merging_utils.heuristic_merging(pairs, patch_id, sim[pairs[:, 0], pairs[:, 1]].astype(np.float64))
        assert np.array_equal(seg, seg_host)
