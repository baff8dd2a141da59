"""GPU parity of patch extraction: cpfn_extract_patches through the C ABI vs the numpy
oracle (indices in stable-argsort order and distances BIT-EXACT), the host loop ``sample`` vs the oracle
under the same np.random seed and vs the goldens of the unmodified reference, tie-breaks, edge sizes, and
size-independent properties at the full sizes of BASELINE configs 3 / 5 (131 072 and 1 M points)."""
import os

import numpy as np
import pytest
import torch

from cpfn_b200 import sampling_utils, synth
from oracle import patches as opatch
from tests.golden import cases

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_patches.npz"))


def _run(hr, seeds, k, dev):
    idx, dist, radius = sampling_utils.extract_patches(torch.from_numpy(hr).to(dev), torch.from_numpy(seeds).to(dev),
                                                       k, return_distances=True)
    return idx.cpu().numpy(), dist.cpu().numpy(), radius.cpu().numpy()


def _check_exact(hr, seeds, k, dev):
    idx, dist, radius = _run(hr, seeds, k, dev)
    k = min(k, hr.shape[0])
    assert idx.shape == (len(seeds), k) and idx.dtype == np.int32
    for s in range(len(seeds)):
        oi, od = opatch.nearest(seeds[s], hr, k)
        assert np.array_equal(idx[s], oi), s
        assert np.array_equal(dist[s].view(np.uint32), od.view(np.uint32)), s
        assert radius[s] == od[-1]


@pytest.mark.parametrize("name", sorted(cases.patch_cases()))
def test_batched_seeds_match_oracle(name, cuda_dev):
    lr, hr, pool, _, k, _, _ = cases.patch_cases()[name]
    _check_exact(hr, np.ascontiguousarray(lr[pool[:7]]), k, cuda_dev)


@pytest.mark.parametrize("N,k,S", [(1, 1, 1), (2, 2, 3), (37, 37, 2), (1000, 1, 4), (2049, 2048, 1), (70000, 16384, 2),
                                   (4097, 300, 33)])
def test_edge_sizes(N, k, S, cuda_dev):
    rng = np.random.RandomState(N + k)
    hr = rng.uniform(-1, 1, (N, 3)).astype(np.float32)
    seeds = rng.uniform(-1, 1, (S, 3)).astype(np.float32)
    _check_exact(hr, seeds, k, cuda_dev)


def test_ties_are_ordered_by_index(cuda_dev):
    """Integer lattice (many exactly equal distances straddling the cut), duplicated points, and a cloud of
    identical points: the result must be the stable argsort, which exercises the index digits of the select."""
    lat = synth.lattice_cloud(1, 4096, seed=5)[0].astype(np.float32)
    seeds = np.ascontiguousarray(lat[[0, 17, 900]])
    for k in (1, 7, 64, 500, 2048):
        _check_exact(lat, seeds, k, cuda_dev)
    dup = np.concatenate([lat[:1500]] * 3)
    _check_exact(dup, seeds, 1000, cuda_dev)
    same = np.full((5000, 3), 0.25, np.float32)
    _check_exact(same, np.zeros((2, 3), np.float32), 100, cuda_dev)
    _check_exact(same, np.zeros((1, 3), np.float32), 5000, cuda_dev)


def test_k_is_clamped_to_the_cloud_and_errors_are_loud(cuda_dev):
    hr = np.random.RandomState(0).randn(100, 3).astype(np.float32)
    idx = sampling_utils.extract_patches(torch.from_numpy(hr).to(cuda_dev), torch.zeros(1, 3, device=cuda_dev), 8192)
    assert idx.shape == (1, 100) and sorted(idx[0].tolist()) == list(range(100))
    big = torch.zeros(20000, 3, device=cuda_dev)
    with pytest.raises(RuntimeError):
        sampling_utils.extract_patches(big, torch.zeros(1, 3, device=cuda_dev), 16385)
    with pytest.raises(RuntimeError):
        sampling_utils.extract_patches(big.cpu(), torch.zeros(1, 3), 16)
    with pytest.raises(RuntimeError):
        sampling_utils.extract_patches(big.double(), torch.zeros(1, 3, device=cuda_dev).double(), 16)
    with pytest.raises(TypeError):
        sampling_utils.sample(hr.astype(np.float64), hr.astype(np.float64), np.arange(10), 16, 2)


@pytest.mark.parametrize("name", sorted(cases.patch_cases()))
def test_sample_loop_matches_oracle_and_reference(name, cuda_dev):
    lr, hr, pool, labels, k, mp, seed = cases.patch_cases()[name]
    np.random.seed(seed)
    want = opatch.sample(lr, hr, pool.copy(), num_points_patch=k, max_number_patches=mp)
    np.random.seed(seed)
    got = sampling_utils.sample(lr, hr, pool.copy(), num_points_patch=k, max_number_patches=mp, device=cuda_dev)
    assert got.dtype == np.int64 and np.array_equal(got, want)
    cases.assert_patches_equivalent(got, GOLD[name + "/sample"], hr)
    np.random.seed(seed)
    want = opatch.sample_per_label(lr, hr, pool.copy(), labels.copy(), num_points_patch=k, max_number_patches=mp)
    np.random.seed(seed)
    got = sampling_utils.sample_per_label(lr, hr, pool.copy(), labels.copy(), num_points_patch=k,
                                          max_number_patches=mp, device=cuda_dev)
    assert np.array_equal(got, want)
    cases.assert_patches_equivalent(got, GOLD[name + "/sample_per_label"], hr)


@pytest.mark.parametrize("N", [131072, 1 << 20])
def test_full_size_properties(N, cuda_dev):
    """BASELINE config 3 / 5 sizes, 32 seeds, k = 8192: rows sorted, indices unique, the selected set is
    exactly {d < radius} plus ties at the radius (counted on the device), two seeds checked against the oracle."""
    hr = synth.shape_cloud(N, 4242)[0].astype(np.float32)
    seeds = np.ascontiguousarray(hr[:: N // 32][:32])
    hr_d, seeds_d = torch.from_numpy(hr).to(cuda_dev), torch.from_numpy(seeds).to(cuda_dev)
    idx, dist, radius = sampling_utils.extract_patches(hr_d, seeds_d, 8192, return_distances=True)
    assert bool((dist[:, 1:] >= dist[:, :-1]).all()) and bool((radius == dist[:, -1]).all())
    srt = torch.sort(idx.long(), dim=1).values
    assert bool((srt[:, 1:] > srt[:, :-1]).all()) and int(srt.min()) >= 0 and int(srt.max()) < N
    for s in (0, 13, 31):
        d_all = (seeds_d[s][None] - hr_d).double().norm(dim=1)           # fp64: only used with a margin
        r = float(radius[s])
        assert int((d_all < r * (1 - 1e-5)).sum()) <= 8192 <= int((d_all <= r * (1 + 1e-5)).sum())
        picked = torch.zeros(N, dtype=torch.bool, device=cuda_dev)
        picked[idx[s].long()] = True
        assert bool(picked[d_all < r * (1 - 1e-5)].all()) and not bool(picked[d_all > r * (1 + 1e-5)].any())
    for s in (3, 29):
        oi, od = opatch.nearest(seeds[s], hr, 8192)
        assert np.array_equal(idx[s].cpu().numpy(), oi) and np.array_equal(dist[s].cpu().numpy(), od)
