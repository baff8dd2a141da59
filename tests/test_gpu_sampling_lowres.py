"""GPU parity of the low-resolution farthest point sampling (cpfn_fps_dense through the C ABI):
indices BIT-EXACT against the numpy oracle and the goldens of the unmodified reference functions, incl. the
first-maximum tie-break on a lattice with duplicated points, seeds, the per-label variant under the same
np.random seed, and the full preprocessing size (131 072 -> 8192) against size-independent properties."""
import os

import numpy as np
import pytest
import torch

from cpfn_b200 import sampling_lowres, synth
from oracle import sampling_lowres as olow
from tests.golden import cases

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_sampling_lowres.npz"))


@pytest.mark.parametrize("name", sorted(cases.lowres_cases()))
def test_matches_reference_and_oracle(name, cuda_dev):
    P, labels, m, seed = cases.lowres_cases()[name]
    np.random.seed(seed)
    first = sampling_lowres.furthest_point_sampling_per_label(P, labels, device=cuda_dev)
    assert first.dtype == np.int32 and np.array_equal(first, GOLD[name + "/per_label"])
    second = sampling_lowres.furthest_point_sampling(P, first, m, device=cuda_dev)
    assert np.array_equal(second, GOLD[name + "/fps"])
    assert np.array_equal(sampling_lowres.furthest_point_sampling(P, np.zeros(0, np.int32), 64, device=cuda_dev),
                          GOLD[name + "/fps_unseeded"])
    np.random.seed(seed)
    both = sampling_lowres.lowres_indices(P, labels, m, device=cuda_dev)
    assert np.array_equal(both, np.concatenate((first, second)))


@pytest.mark.parametrize("N,m", [(1, 1), (5, 5), (1025, 40), (70000, 100), (300000, 64)])
def test_sizes_against_oracle(N, m, cuda_dev):
    rng = np.random.RandomState(N)
    P = rng.uniform(-1, 1, (N, 3)).astype(np.float32)
    seeds = rng.choice(N, min(3, N), replace=False).astype(np.int32)
    assert np.array_equal(sampling_lowres.furthest_point_sampling(P, seeds, m, device=cuda_dev),
                          olow.furthest_point_sampling(P, seeds, m))


def test_errors_are_loud(cuda_dev):
    with pytest.raises(TypeError):
        sampling_lowres.furthest_point_sampling(np.zeros((10, 3)), np.zeros(0, np.int32), 2, device=cuda_dev)
    with pytest.raises(RuntimeError):                    # more points than the register-resident kernel holds
        sampling_lowres.furthest_point_sampling(np.zeros((3_000_000, 3), np.float32), np.zeros(0, np.int32), 2, device=cuda_dev)


def test_full_preprocessing_size(cuda_dev):
    """131 072 points -> per-label seeds + 8192 samples (the script's default): distinct indices, every label
    present, the running minimum never increases along the sample order, and the first 300 samples agree with
    the oracle run on the same cloud."""
    N, m = 131072, 8192
    P, _, _, I = synth.shape_batch(1, N, seed=61)
    P, labels = P[0].astype(np.float32), I[0].astype(np.int32)
    np.random.seed(7)
    idx = sampling_lowres.lowres_indices(P, labels, m, device=cuda_dev)
    n_lab = len(np.unique(labels))
    assert idx.shape == (n_lab + m,) and len(np.unique(labels[idx[:n_lab]])) == n_lab
    second = idx[n_lab:]
    assert len(np.unique(second)) == m and not set(second) & set(idx[:n_lab])
    Pd = torch.from_numpy(P).to(cuda_dev)
    chosen = Pd[torch.from_numpy(second.astype(np.int64)).to(cuda_dev)]
    # distance of sample i to the samples before it (seeds do not count: they only start at distance 0)
    d = torch.cdist(chosen[:2048].double(), chosen[:2048].double())
    prev_min = torch.stack([d[i, :i].min() for i in range(1, 2048)])
    assert bool((prev_min[1:] <= prev_min[:-1] + 1e-6).all())
    np.random.seed(7)
    o_first = olow.furthest_point_sampling_per_label(P, labels)
    assert np.array_equal(idx[:n_lab], o_first)
    assert np.array_equal(second[:300], olow.furthest_point_sampling(P, o_first, 300))
