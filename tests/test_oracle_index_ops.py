"""CPU checks of the C oracle (oracle/cpfn_oracle.c) against an independent numpy
statement of the reference semantics.

Inputs are dyadic lattices (coordinates k/32): every product and sum is exact in
fp32, so fused and unfused arithmetic agree and plain numpy float32 is a valid
second implementation.  The closed-form FPS tie-break (bit-reversed thread id,
then lowest k) is checked against the oracle's literal shared-memory-tree
simulation.
"""
import numpy as np
import pytest

from cpfn_b200 import synth


def _bitrev(v, bits):
    r = 0
    for i in range(bits):
        r |= ((v >> i) & 1) << (bits - 1 - i)
    return r


def _fps_numpy(xyz, m):
    n = xyz.shape[0]
    log2t = min(9, int(np.floor(np.log2(n))))
    T = 1 << log2t
    rank = np.array([(_bitrev(k % T, log2t), k // T) for k in range(n)])
    order = np.lexsort((rank[:, 1], rank[:, 0]))  # best rank first
    pos = np.empty(n, dtype=np.int64)
    pos[order] = np.arange(n)
    mag = (xyz.astype(np.float32) ** 2).sum(1, dtype=np.float32)
    live = ~(mag.astype(np.float64) <= 1e-3)
    temp = np.full(n, 1e10, dtype=np.float32)
    out = np.zeros(m, dtype=np.int32)
    old = 0
    for j in range(1, m):
        d = ((xyz - xyz[old]) ** 2).sum(1, dtype=np.float32)
        temp = np.where(live, np.minimum(d, temp), temp)
        cand = np.where(live)[0]
        if cand.size == 0:
            old = 0
        else:
            best = temp[cand].max()
            tied = cand[temp[cand] == best]
            old = int(tied[np.argmin(pos[tied])])
        out[j] = old
    return out


def _dyadic(batch, n, seed, span=32):
    rng = np.random.default_rng(seed)
    return (rng.integers(-span, span + 1, size=(batch, n, 3)) / 32.0).astype(np.float32)


@pytest.mark.parametrize("n,m", [(1, 1), (2, 2), (5, 5), (31, 16), (33, 20), (100, 64), (259, 128),
                                 (512, 128), (700, 200), (1500, 256)])
def test_fps_matches_closed_form_tie_break(oracle_ops, n, m):
    xyz = _dyadic(3, n, seed=n)
    got = oracle_ops.farthest_point_sampling(xyz, m)
    for b in range(xyz.shape[0]):
        np.testing.assert_array_equal(got[b], _fps_numpy(xyz[b], m))


def test_fps_m_larger_than_distinct_points(oracle_ops):
    xyz = np.tile(_dyadic(1, 4, seed=3), (1, 8, 1))  # 32 points, 4 distinct
    got = oracle_ops.farthest_point_sampling(xyz, 20)
    np.testing.assert_array_equal(got[0], _fps_numpy(xyz[0], 20))


def test_fps_skip_rule_boundary(oracle_ops):
    # |p|^2 <= 1e-3 (double compare): 0x3A83126E (< 0.001f) is skipped, 0.001f is kept.
    below = np.array([0x3A83126E], dtype=np.uint32).view(np.float32)[0]
    at = np.float32(0.001)
    xyz = np.zeros((1, 64, 3), dtype=np.float32)
    xyz[0, :, 0] = np.linspace(0.5, 1.0, 64, dtype=np.float32)
    xyz[0, 10] = [np.sqrt(below), 0, 0]
    xyz[0, 11] = [np.sqrt(at) * np.float32(1.0001), 0, 0]
    got = oracle_ops.farthest_point_sampling(xyz, 64)[0]
    mag = (xyz[0, :, 0] * xyz[0, :, 0]).astype(np.float32)
    skipped = set(np.where(mag.astype(np.float64) <= 1e-3)[0].tolist())
    assert 10 in skipped and 11 not in skipped
    assert not (set(got[1:].tolist()) & skipped)
    assert 11 in set(got.tolist())


def test_fps_all_points_skipped_returns_zero(oracle_ops):
    xyz = np.full((2, 40, 3), 0.001, dtype=np.float32)
    assert (oracle_ops.farthest_point_sampling(xyz, 8) == 0).all()


def test_opt_n_threads_matches_survey(oracle_ops):
    assert oracle_ops.opt_n_threads(8192) == 512
    assert oracle_ops.opt_n_threads(512) == 512
    assert oracle_ops.opt_n_threads(259) == 256
    assert oracle_ops.opt_n_threads(1) == 1
    for p in range(0, 10):
        assert oracle_ops.opt_n_threads(1 << p) == (1 << p)


def _ball_numpy(q, xyz, radius, ns):
    r2 = np.float32(radius) * np.float32(radius)
    out = np.zeros((q.shape[0], ns), dtype=np.int32)
    for j in range(q.shape[0]):
        d = ((q[j] - xyz) ** 2).sum(1, dtype=np.float32)
        hits = np.where(d < r2)[0][:ns]
        if hits.size:
            out[j] = hits[0]
            out[j, :hits.size] = hits
    return out


@pytest.mark.parametrize("n,s,radius,ns", [(64, 8, 0.25, 4), (500, 33, 0.5, 64), (2000, 100, 0.125, 16),
                                           (300, 20, 1e-4, 8), (1, 1, 1.0, 3)])
def test_ball_query_first_k_in_index_order(oracle_ops, n, s, radius, ns):
    xyz = _dyadic(2, n, seed=7 * n)
    q = xyz[:, :s].copy()
    q[:, -1] += 100.0  # one query with an empty ball -> all zeros
    got = oracle_ops.ball_query(q, xyz, radius, ns)
    for b in range(2):
        np.testing.assert_array_equal(got[b], _ball_numpy(q[b], xyz[b], radius, ns))
    assert (got[:, -1] == 0).all()


def _three_nn_numpy(u, kn):
    d = ((u[:, None, :] - kn[None, :, :]) ** 2).sum(2, dtype=np.float32)
    idx = np.argsort(d, axis=1, kind="stable")[:, :3]  # stable: lower index wins ties
    return np.take_along_axis(d, idx, 1), idx.astype(np.int32)


@pytest.mark.parametrize("n,m", [(50, 3), (200, 17), (1000, 128)])
def test_three_nn_ties_go_to_lower_index(oracle_ops, n, m):
    u = _dyadic(2, n, seed=n + 1, span=8)
    kn = _dyadic(2, m, seed=m + 2, span=8)  # coarse lattice: many exact ties
    d2, idx = oracle_ops.three_nn(u, kn)
    for b in range(2):
        rd, ri = _three_nn_numpy(u[b], kn[b])
        np.testing.assert_array_equal(idx[b], ri)
        np.testing.assert_array_equal(d2[b], rd)


def test_three_nn_fewer_than_three_known(oracle_ops):
    u = _dyadic(1, 10, seed=1)
    d2, idx = oracle_ops.three_nn(u, u[:, :2])
    assert np.isinf(d2[..., 2]).all() and (idx[..., 2] == 0).all()
    assert np.isfinite(d2[..., :2]).all()


def test_weighted_sum_gather_group_and_grads(oracle_ops):
    rng = np.random.default_rng(5)
    B, C, M, n, S, K = 2, 5, 40, 30, 6, 4
    pts = rng.normal(size=(B, C, M)).astype(np.float32)
    idx3 = rng.integers(0, M, size=(B, n, 3)).astype(np.int32)
    w = rng.random(size=(B, n, 3)).astype(np.float32)
    out = oracle_ops.three_weighted_sum(pts, idx3, w)
    ref = sum(np.take_along_axis(pts, np.broadcast_to(idx3[:, None, :, i], (B, C, n)), 2).astype(np.float64)
              * w[:, None, :, i] for i in range(3))
    np.testing.assert_allclose(out, ref, rtol=1e-6, atol=1e-6)
    g = rng.normal(size=(B, C, n)).astype(np.float32)
    gp = oracle_ops.three_weighted_sum_grad(g, idx3, w, M)
    # adjoint identity <g, A p> == <A^T g, p>
    np.testing.assert_allclose((g.astype(np.float64) * out).sum(), (gp.astype(np.float64) * pts).sum(),
                               rtol=1e-4)
    idx = rng.integers(0, M, size=(B, n)).astype(np.int32)
    ga = oracle_ops.gather_points(pts, idx)
    np.testing.assert_array_equal(ga, np.take_along_axis(pts, np.broadcast_to(idx[:, None], (B, C, n)), 2))
    gg = oracle_ops.gather_points_grad(g, idx, M)
    np.testing.assert_allclose((g.astype(np.float64) * ga).sum(), (gg.astype(np.float64) * pts).sum(), rtol=1e-4)
    gidx = rng.integers(0, M, size=(B, S, K)).astype(np.int32)
    gr = oracle_ops.group_points(pts, gidx)
    assert gr.shape == (B, C, S, K)
    np.testing.assert_array_equal(gr.reshape(B, C, S * K),
                                  oracle_ops.gather_points(pts, gidx.reshape(B, S * K)))
    go = rng.normal(size=(B, C, S, K)).astype(np.float32)
    grg = oracle_ops.group_points_grad(go, gidx, M)
    np.testing.assert_allclose((go.astype(np.float64) * gr).sum(), (grg.astype(np.float64) * pts).sum(), rtol=1e-4)


def test_synthetic_clouds_are_deterministic_and_normalised():
    P, X, W, I = synth.shape_batch(2, 1024, seed=3, k_slots=24)
    P2, _, W2, _ = synth.shape_batch(2, 1024, seed=3, k_slots=24)
    np.testing.assert_array_equal(P, P2)
    np.testing.assert_array_equal(W, W2)
    assert P.dtype == np.float32 and W.shape == (2, 1024, 24)
    np.testing.assert_allclose(np.linalg.norm(P, axis=2).max(axis=1), 1.0, rtol=1e-6)
    np.testing.assert_allclose(np.linalg.norm(X, axis=2), 1.0, rtol=1e-5)
    np.testing.assert_allclose(W.sum(2), 1.0, rtol=1e-5)
    L = synth.lattice_cloud(1, 256, seed=1)
    assert (np.einsum("bnc,bnc->bn", L, L) <= 1e-3).any()
