"""GPU parity of the nine pointnet2 ops: CUDA path (through the C ABI) vs the C
oracle, vs the committed golden vectors of the reference kernels, and -- when
oracle/_ref is present -- vs the reference kernels live on the same device.
Integer outputs and three_nn distances must be bit-exact."""
import os

import numpy as np
import pytest
import torch

from cpfn_b200 import cuda_ops, synth
from tests.golden import cases

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ref_cuda_ops.npz")


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def _fps(dev):
    return lambda xyz, m: cuda_ops.farthest_point_sampling(_t(xyz, dev), m).cpu().numpy()


@pytest.fixture(scope="module")
def golden():
    if not os.path.exists(GOLDEN):
        pytest.skip("golden vectors not generated yet")
    return np.load(GOLDEN)


@pytest.fixture(scope="module")
def ref_ext():
    from oracle import build_ref
    mod = build_ref.load_module()
    if mod is None:
        pytest.skip("oracle/_ref/ref_cuda_ops.so not present")
    return mod


FPS_SIZES = [(1, 1), (2, 2), (7, 7), (31, 31), (32, 8), (33, 33), (63, 40), (64, 64), (100, 64),
             (259, 128), (511, 100), (512, 128), (513, 128), (1000, 300), (1023, 64), (1024, 256),
             (1025, 64), (2048, 128), (3000, 100), (4096, 512), (5000, 64), (8192, 512),
             (8193, 32), (12000, 64), (16384, 32), (16385, 24), (40000, 16)]


@pytest.mark.parametrize("n,m", FPS_SIZES)
def test_fps_bit_exact_vs_oracle(cuda_dev, oracle_ops, n, m):
    for kind in ("uniform", "lattice"):
        xyz = (synth.uniform_cloud(3, n, seed=n) if kind == "uniform"
               else synth.lattice_cloud(3, n, seed=n, pitch=8))
        got = cuda_ops.farthest_point_sampling(_t(xyz, cuda_dev), m).cpu().numpy()
        np.testing.assert_array_equal(got, oracle_ops.farthest_point_sampling(xyz, m), err_msg=kind)


def test_fps_full_size_batch16(cuda_dev, oracle_ops):
    P = synth.shape_batch(16, 8192, seed=1235)[0]
    got = cuda_ops.farthest_point_sampling(_t(P, cuda_dev), 512).cpu().numpy()
    np.testing.assert_array_equal(got, oracle_ops.farthest_point_sampling(P, 512))
    # size-independent property: indices are distinct and start at 0
    assert (got[:, 0] == 0).all()
    assert all(len(set(r.tolist())) == 512 for r in got)


def test_fps_skip_rule_and_degenerate(cuda_dev, oracle_ops):
    xyz = np.full((2, 100, 3), 0.001, dtype=np.float32)  # everything skipped
    assert (cuda_ops.farthest_point_sampling(_t(xyz, cuda_dev), 9).cpu().numpy() == 0).all()
    xyz = synth.uniform_cloud(2, 2000, seed=5) * np.float32(0.05)  # most points inside the skip ball
    got = cuda_ops.farthest_point_sampling(_t(xyz, cuda_dev), 200).cpu().numpy()
    np.testing.assert_array_equal(got, oracle_ops.farthest_point_sampling(xyz, 200))
    dup = np.tile(synth.uniform_cloud(1, 4, seed=1), (1, 16, 1))  # m > distinct points
    got = cuda_ops.farthest_point_sampling(_t(dup, cuda_dev), 30).cpu().numpy()
    np.testing.assert_array_equal(got, oracle_ops.farthest_point_sampling(dup, 30))


def test_fps_empty(cuda_dev):
    out = cuda_ops.farthest_point_sampling(torch.zeros(0, 16, 3, device=cuda_dev), 4)
    assert out.shape == (0, 4)
    out = cuda_ops.farthest_point_sampling(torch.zeros(2, 16, 3, device=cuda_dev), 0)
    assert out.shape == (2, 0)


def test_ball_query_bit_exact_vs_oracle(cuda_dev, oracle_ops):
    for name, (q, xyz, r, k) in cases.ball_cases(oracle_ops.farthest_point_sampling).items():
        got = cuda_ops.ball_query(_t(q, cuda_dev), _t(xyz, cuda_dev), r, k).cpu().numpy()
        np.testing.assert_array_equal(got, oracle_ops.ball_query(q, xyz, r, k), err_msg=name)


@pytest.mark.parametrize("n,s,k,r", [(1, 1, 1, 1.0), (5, 5, 8, 0.5), (33, 7, 3, 0.3), (700, 129, 17, 0.2),
                                     (8192, 512, 64, 0.2), (9000, 40, 64, 0.1), (20000, 300, 128, 0.08),
                                     (30000, 16, 64, 0.01)])
def test_ball_query_sizes(cuda_dev, oracle_ops, n, s, k, r):
    xyz = synth.uniform_cloud(2, n, seed=n + s)
    q = xyz[:, ::max(1, n // s)][:, :s].copy()
    got = cuda_ops.ball_query(_t(q, cuda_dev), _t(xyz, cuda_dev), r, k).cpu().numpy()
    np.testing.assert_array_equal(got, oracle_ops.ball_query(q, xyz, r, k))


def test_ball_query_full_size_property(cuda_dev, oracle_ops):
    P = synth.shape_batch(16, 8192, seed=1235)[0]
    Pd = _t(P, cuda_dev)
    fidx = cuda_ops.farthest_point_sampling(Pd, 512)
    q = torch.gather(Pd, 1, fidx.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    got = cuda_ops.ball_query(q, Pd, 0.2, 64)
    np.testing.assert_array_equal(got.cpu().numpy(), oracle_ops.ball_query(q.cpu().numpy(), P, 0.2, 64))
    # properties: every returned point is inside the ball; rows ascend until the padding starts
    nb = torch.gather(Pd.unsqueeze(1).expand(-1, 512, -1, -1), 2,
                      got.long().unsqueeze(-1).expand(-1, -1, -1, 3))
    assert (((nb - q.unsqueeze(2)) ** 2).sum(-1) < 0.2 * 0.2 + 1e-6).all()


def test_three_nn_bit_exact_vs_oracle(cuda_dev, oracle_ops):
    for name, (u, kn) in cases.three_nn_cases(oracle_ops.farthest_point_sampling).items():
        d2, idx = cuda_ops.three_nn(_t(u, cuda_dev), _t(kn, cuda_dev))
        rd, ri = oracle_ops.three_nn(u, kn)
        np.testing.assert_array_equal(idx.cpu().numpy(), ri, err_msg=name)
        np.testing.assert_array_equal(d2.cpu().numpy(), rd, err_msg=name)


@pytest.mark.parametrize("n,m", [(1, 1), (10, 2), (300, 3), (8192, 512), (1000, 2049), (513, 5000)])
def test_three_nn_sizes(cuda_dev, oracle_ops, n, m):
    u = synth.uniform_cloud(2, n, seed=n)
    kn = synth.lattice_cloud(2, m, seed=m, pitch=10)
    d2, idx = cuda_ops.three_nn(_t(u, cuda_dev), _t(kn, cuda_dev))
    rd, ri = oracle_ops.three_nn(u, kn)
    np.testing.assert_array_equal(idx.cpu().numpy(), ri)
    np.testing.assert_array_equal(d2.cpu().numpy(), rd)


def test_weighted_sum_gather_group_vs_oracle(cuda_dev, oracle_ops):
    z = cases.interp_inputs()
    pts, idx, w, g = (_t(z[k], cuda_dev) for k in ("pts", "idx", "w", "g"))
    np.testing.assert_array_equal(cuda_ops.three_weighted_sum(pts, idx, w).cpu().numpy(),
                                  oracle_ops.three_weighted_sum(z["pts"], z["idx"], z["w"]))
    # scatter-adds: float atomics, order differs -> tolerance 1e-5 relative to the row scale
    np.testing.assert_allclose(cuda_ops.three_weighted_sum_grad(g, idx, w, z["M"]).cpu().numpy(),
                               oracle_ops.three_weighted_sum_grad(z["g"], z["idx"], z["w"], z["M"]),
                               rtol=1e-5, atol=1e-5)
    gi = np.ascontiguousarray(z["idx"][:, :, 0])
    np.testing.assert_array_equal(cuda_ops.gather_points(pts, _t(gi, cuda_dev)).cpu().numpy(),
                                  oracle_ops.gather_points(z["pts"], gi))
    np.testing.assert_allclose(cuda_ops.gather_points_grad(g, _t(gi, cuda_dev), z["M"]).cpu().numpy(),
                               oracle_ops.gather_points_grad(z["g"], gi, z["M"]), rtol=1e-5, atol=1e-5)
    gidx, gg = _t(z["gidx"], cuda_dev), _t(z["gg"], cuda_dev)
    np.testing.assert_array_equal(cuda_ops.group_points(pts, gidx).cpu().numpy(),
                                  oracle_ops.group_points(z["pts"], z["gidx"]))
    np.testing.assert_allclose(cuda_ops.group_points_grad(gg, gidx, z["M"]).cpu().numpy(),
                               oracle_ops.group_points_grad(z["gg"], z["gidx"], z["M"]),
                               rtol=1e-5, atol=1e-5)


def test_against_golden_reference_outputs(cuda_dev, golden):
    for name, (xyz, m) in cases.fps_cases().items():
        got = cuda_ops.farthest_point_sampling(_t(xyz, cuda_dev), m).cpu().numpy()
        np.testing.assert_array_equal(got, golden["fps/" + name].astype(np.int32), err_msg=name)
    for name, (q, xyz, r, k) in cases.ball_cases(_fps(cuda_dev)).items():
        got = cuda_ops.ball_query(_t(q, cuda_dev), _t(xyz, cuda_dev), r, k).cpu().numpy()
        np.testing.assert_array_equal(got, golden["ball/" + name].astype(np.int32), err_msg=name)
    for name, (u, kn) in cases.three_nn_cases(_fps(cuda_dev)).items():
        d2, idx = cuda_ops.three_nn(_t(u, cuda_dev), _t(kn, cuda_dev))
        np.testing.assert_array_equal(idx.cpu().numpy(), golden["nn_idx/" + name].astype(np.int32))
        np.testing.assert_array_equal(d2.cpu().numpy(), golden["nn_d2/" + name])
    z = cases.interp_inputs()
    pts, idx, w = (_t(z[k], cuda_dev) for k in ("pts", "idx", "w"))
    np.testing.assert_array_equal(cuda_ops.three_weighted_sum(pts, idx, w).cpu().numpy(), golden["tws"])


def test_live_against_reference_extension(cuda_dev, ref_ext):
    """Same device, same inputs, the UNMODIFIED reference kernels vs ours."""
    P = _t(synth.shape_batch(16, 8192, seed=1235)[0], cuda_dev)
    a = cuda_ops.farthest_point_sampling(P, 512)
    b = ref_ext.farthest_point_sampling(P, 512)
    assert torch.equal(a, b)
    q = torch.gather(P, 1, a.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    assert torch.equal(cuda_ops.ball_query(q, P, 0.2, 64), ref_ext.ball_query(q, P, 0.2, 64))
    d2a, ia = cuda_ops.three_nn(P, q)
    d2b, ib = ref_ext.three_nn(P, q)
    assert torch.equal(ia, ib) and torch.equal(d2a, d2b)
    feats = torch.randn(16, 32, 512, device=cuda_dev)
    w = torch.rand(16, 8192, 3, device=cuda_dev)
    assert torch.equal(cuda_ops.three_weighted_sum(feats, ia, w), ref_ext.three_weighted_sum(feats, ib, w))
    g = torch.randn(16, 32, 8192, device=cuda_dev)
    torch.testing.assert_close(cuda_ops.three_weighted_sum_grad(g, ia, w, 512),
                               ref_ext.three_weighted_sum_grad(g, ib, w, 512), rtol=1e-4, atol=1e-4)
    gi = cuda_ops.ball_query(q, P, 0.2, 64)
    f2 = torch.randn(16, 8, 8192, device=cuda_dev)
    assert torch.equal(cuda_ops.group_points(f2, gi), ref_ext.group_points(f2, gi))
    assert torch.equal(cuda_ops.gather_points(f2, a), ref_ext.gather_points(f2, a))
    lat = _t(synth.lattice_cloud(4, 5000, seed=9, pitch=12), cuda_dev)
    assert torch.equal(cuda_ops.farthest_point_sampling(lat, 700), ref_ext.farthest_point_sampling(lat, 700))


def test_three_nn_weights_and_gather_xyz_and_post(cuda_dev, oracle_ops):
    """Fused helpers (csrc/glue.cu) against the op sequences they replace."""
    from cpfn_b200 import fused
    rng = np.random.default_rng(7)
    u = synth.uniform_cloud(3, 1000, seed=3)
    k = synth.lattice_cloud(3, 77, seed=4, pitch=8)
    w, idx = fused.three_nn_weights(_t(u, cuda_dev), _t(k, cuda_dev))
    d2, ridx = oracle_ops.three_nn(u, k)
    np.testing.assert_array_equal(idx.cpu().numpy(), ridx)
    d = torch.sqrt(torch.from_numpy(d2))
    r = 1.0 / (d + 1e-8)
    ref_w = (r / torch.sum(r, dim=2, keepdim=True)).numpy()
    np.testing.assert_allclose(w.cpu().numpy(), ref_w, rtol=3e-7, atol=0)
    fi = rng.integers(0, 1000, size=(3, 50)).astype(np.int32)
    g = fused.gather_xyz(_t(u, cuda_dev), _t(fi, cuda_dev)).cpu().numpy()
    np.testing.assert_array_equal(g, np.take_along_axis(u, fi.astype(np.int64)[:, :, None], axis=1))
    h = rng.normal(size=(2, 333, 35)).astype(np.float32) * 3
    X, W, inst, typ = fused.spfn_post(_t(h, cuda_dev), 0, 7, 28, t_off=3, n_types=4)
    np.testing.assert_array_equal(inst.cpu().numpy(), h[:, :, 7:].argmax(2))
    np.testing.assert_array_equal(typ.cpu().numpy(), h[:, :, 3:7].argmax(2))
    ht = torch.from_numpy(h)
    np.testing.assert_allclose(X.cpu().numpy(), torch.nn.functional.normalize(ht[:, :, :3], dim=2).numpy(), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(W.cpu().numpy(), torch.softmax(ht[:, :, 7:], dim=2).numpy(), rtol=2e-6, atol=1e-9)


def test_ball_query_grid_equals_scan_kernel(cuda_dev, oracle_ops, monkeypatch):
    """The uniform-grid kernel and the brute-force scan kernel give the same indices (and both equal
    the oracle) on uniform, lattice (ties on the radius), clustered and out-of-box query sets."""
    rng = np.random.default_rng(12)
    for name, (N, S, r, K) in {"u": (8192, 512, 0.2, 64), "l": (3000, 300, 0.25, 16), "c": (5000, 128, 0.05, 32),
                               "big_r": (2048, 64, 1.5, 64), "k1": (2700, 90, 0.3, 1), "max": (32768, 40, 0.1, 8)}.items():
        xyz = synth.lattice_cloud(2, N, seed=N, pitch=8) if name == "l" else synth.uniform_cloud(2, N, seed=N)
        if name == "c":
            xyz = (xyz * 0.1 + rng.normal(scale=0.3, size=(2, 1, 3))).astype(np.float32)
        q = xyz[:, rng.permutation(N)[:S]].copy()
        q[:, :5] += np.float32(3.0)                       # queries far outside the bounding box: no hits
        q[:, 5:10] += np.float32(r * 0.9)                 # just outside / near the boundary cells
        monkeypatch.delenv("CPFN_BQ_NO_GRID", raising=False)
        a = cuda_ops.ball_query(_t(q, cuda_dev), _t(xyz, cuda_dev), r, K).cpu().numpy()
        monkeypatch.setenv("CPFN_BQ_NO_GRID", "1")
        b = cuda_ops.ball_query(_t(q, cuda_dev), _t(xyz, cuda_dev), r, K).cpu().numpy()
        np.testing.assert_array_equal(a, b, err_msg=name)
        np.testing.assert_array_equal(a, oracle_ops.ball_query(q, xyz, r, K), err_msg=name)


@pytest.mark.parametrize("B,N,m", [(16, 8192, 512), (2, 8192, 300), (3, 1000, 64), (40, 2048, 128), (2, 20000, 96)])
def test_fps_writes_the_centroids(B, N, m, cuda_dev):
    """return_centroids: the sampling kernels (cluster, single-CTA and streaming variants) write xyz[idx] themselves;
    same indices as the plain call, coordinates bit-identical to a gather."""
    P = _t(synth.uniform_cloud(B, N, seed=N + m), cuda_dev)
    idx = cuda_ops.farthest_point_sampling(P, m)
    idx2, cen = cuda_ops.farthest_point_sampling(P, m, return_centroids=True)
    assert torch.equal(idx, idx2) and cen.shape == (B, m, 3)
    want = torch.gather(P, 1, idx.long().unsqueeze(2).expand(B, m, 3))
    assert torch.equal(cen, want)


@pytest.mark.parametrize("name", ["shape", "lattice", "outside", "clustered", "m2048"])
def test_three_nn_grid_equals_scan(cuda_dev, oracle_ops, name):
    """The shared-memory grid search (csrc/nn_grid.cuh) against the exhaustive scan and the C oracle: indices and
    squared distances bit for bit -- lattice clouds with many exactly equal distances (the smaller index must win),
    queries far outside the known cloud's bounding box, known points piled up in a few cells."""
    import os
    from cpfn_b200 import cuda_ops, fused, synth
    rng = np.random.default_rng(5)
    if name == "shape":
        u = synth.shape_batch(3, 8192, seed=3)[0]
        k = u[:, :512].copy()
    elif name == "lattice":
        u = synth.lattice_cloud(2, 3000, seed=4, pitch=6)
        k = u[:, :500].copy()
    elif name == "outside":
        k = synth.uniform_cloud(2, 600, seed=5) * np.float32(0.2)
        u = synth.uniform_cloud(2, 2000, seed=6) * np.float32(3.0)
    elif name == "clustered":
        k = (rng.normal(size=(2, 450, 3)) * 0.01).astype(np.float32)
        k[:, :5] += 1.0
        u = synth.uniform_cloud(2, 1500, seed=7)
    else:
        u = synth.uniform_cloud(1, 5000, seed=8)
        k = synth.uniform_cloud(1, 2048, seed=9)
    U, Kn = torch.from_numpy(u).to(cuda_dev), torch.from_numpy(k).to(cuda_dev)
    d_grid, i_grid = cuda_ops.three_nn(U, Kn)
    w_grid, iw_grid = fused.three_nn_weights(U, Kn)
    os.environ["CPFN_NN_NO_GRID"] = "1"
    try:
        d_scan, i_scan = cuda_ops.three_nn(U, Kn)
        w_scan, iw_scan = fused.three_nn_weights(U, Kn)
    finally:
        os.environ.pop("CPFN_NN_NO_GRID")
    assert torch.equal(i_grid, i_scan) and torch.equal(d_grid, d_scan)
    assert torch.equal(iw_grid, iw_scan) and torch.equal(w_grid, w_scan)
    d_ref, i_ref = oracle_ops.three_nn(u, k)
    assert np.array_equal(i_grid.cpu().numpy(), i_ref) and np.array_equal(d_grid.cpu().numpy(), d_ref)


@pytest.mark.parametrize("B,n,m,radius", [(3, 8192, 512, 0.2), (2, 5000, 384, 0.1), (1, 2048, 2048, 0.3)])
def test_three_nn_weights_on_cell_sorted_queries(cuda_dev, B, n, m, radius):
    """cpfn_three_nn_weights_sorted: the queries taken from the cell-ordered copy of the cloud that the ball-query grid
    holds (coherent warps) -- weights and indices identical, query by query, to the kernel on the original order."""
    from cpfn_b200 import _lib, fused, synth
    P = torch.from_numpy(synth.shape_batch(B, n, seed=n + m)[0]).to(cuda_dev)
    known = fused.gather_xyz(P, cuda_ops.farthest_point_sampling(P, m))
    L = _lib.lib()
    nbytes = L.cpfn_ball_query_grid_workspace_bytes(B, n)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=cuda_dev)
    _lib.check(L.cpfn_ball_query_grid_build(P.data_ptr(), B, n, radius, ws.data_ptr(), nbytes,
                                            torch.cuda.current_stream(cuda_dev).cuda_stream), "grid_build")
    w0, i0 = fused.three_nn_weights(P, known)
    w1, i1 = fused.three_nn_weights(P, known, sorted_queries=ws)
    assert torch.equal(i0, i1) and torch.equal(w0, w1)
    recs = ws[: B * n * 16].view(torch.float32).reshape(B, n, 4)
    order = recs[:, :, 3].contiguous().view(torch.int32).long()
    assert torch.equal(order.sort(dim=1)[0], torch.arange(n, device=cuda_dev).expand(B, n))     # a permutation per cloud
    assert torch.equal(torch.gather(P, 1, order.unsqueeze(-1).expand(-1, -1, 3)), recs[:, :, :3])

