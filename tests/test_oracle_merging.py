"""oracle/merging.py against the goldens of the unmodified reference, and the host-side greedy merge of the
library (C, no GPU needed) against the oracle's literal repeated-argmax loop."""
import os

import numpy as np
import pytest

from oracle import merging as omerge
from tests.golden import cases

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_merging.npz"))
CASES = cases.merging_cases()


@pytest.mark.parametrize("name", sorted(CASES))
def test_similarity_and_labels_match_reference(name):
    c = CASES[name]
    nb, Np, Kl = c["W"].shape
    Kg = c["S"].shape[1]
    sim = omerge.similarity_soft(c["S"], c["W"], c["idx"])
    ref = GOLD[name + "/similarity"]
    exact = omerge.similarity_soft(c["S"], c["W"], c["idx"], dtype=np.float64)
    scale = np.abs(exact).max()
    assert np.abs(sim - ref).max() <= 1e-5 * scale          # two fp32 matmuls, different summation orders
    assert np.abs(ref - exact).max() <= 1e-5 * scale
    labels = omerge.run_heuristic_solver(ref, nb, Kg, Kl)   # same matrix in -> same labels out
    assert np.array_equal(labels, GOLD[name + "/labels"])
    fused = omerge.fuse_patches(c["S"], c["W"], c["idx"], labels)
    assert fused.shape == GOLD[name + "/fused"].shape
    assert np.abs(fused - GOLD[name + "/fused"]).max() <= 1e-6


@pytest.mark.parametrize("name", sorted(CASES))
def test_host_solver_matches_literal_loop(name, built_lib):
    from cpfn_b200 import merging_utils
    c = CASES[name]
    nb, Np, Kl = c["W"].shape
    Kg = c["S"].shape[1]
    ref = GOLD[name + "/similarity"]
    assert np.array_equal(merging_utils.run_heuristic_solver(ref, nb, Kg, Kl), GOLD[name + "/labels"])
    for thr in (0.5, 5.0):                                  # thresholds change which pairs exist and which nodes are dropped
        assert np.array_equal(merging_utils.run_heuristic_solver(ref, nb, Kg, Kl, threshold=thr),
                              omerge.run_heuristic_solver(ref, nb, Kg, Kl, threshold=thr))


def test_host_solver_ties_first_pair_and_many_patches(built_lib):
    """Random pair lists with heavily tied penalties (argmax takes the FIRST maximum), a first pair inside one
    patch (the reference merges it before any filtering) and more than 64 patches (multi-word patch sets)."""
    from cpfn_b200 import merging_utils
    rng = np.random.RandomState(0)
    for trial in range(30):
        n_patch = int(rng.choice([2, 3, 7, 70]))
        per = int(rng.randint(1, 5))
        patch_id = np.repeat(np.arange(n_patch), per)
        n = len(patch_id)
        P = int(rng.randint(1, 4 * n))
        a = rng.randint(0, n, P)
        b = rng.randint(0, n, P)
        lo, hi = np.minimum(a, b), np.maximum(a, b)
        keep = lo < hi
        pairs = np.stack((lo[keep], hi[keep]), axis=1).astype(np.int64)
        if len(pairs) == 0:
            continue
        pen = rng.randint(1, 4, len(pairs)).astype(np.float64)       # only three distinct values: ties everywhere
        if trial % 3 == 0 and per > 1:
            pairs[0] = (0, 1)                                          # same patch ...
            pen[0] = 10.0                                              # ... and the global maximum
        want = omerge.heuristic_merging(pairs.copy(), patch_id.copy(), pen.copy())
        got = merging_utils.heuristic_merging(pairs, patch_id, pen)
        assert np.array_equal(got, want), trial
    assert np.array_equal(merging_utils.heuristic_merging(np.zeros((0, 2), np.int64), np.arange(4), np.zeros(0)),
                          np.arange(4))


def test_normals_types_block():
    c = CASES["small"]
    Xg, Tg = omerge.merge_normals_types(c["X"], c["T"], c["idx"], c["obj_normals"], c["obj_types"])
    covered = np.zeros(len(c["S"]), bool)
    covered[c["idx"].reshape(-1)] = True
    assert np.allclose(np.linalg.norm(Xg, axis=1), 1.0, atol=1e-5)
    assert np.array_equal(Tg[~covered], c["obj_types"][~covered]) and 0 < covered.sum() < len(covered)
