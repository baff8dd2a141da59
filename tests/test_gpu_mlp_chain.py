"""GPU parity of the fused tcgen05 MLP-chain kernel (cpfn_mlp_chain) against a plain PyTorch
fp32 reference of the same op (gather / interpolate -> (W x + b, ReLU) x n -> max-pool), on
random data, for every input / output mode, tile size and the channel paddings the network uses.

Tolerance: operands are split-bf16 (hi + lo, 16 mantissa bits; three MMAs per product),
accumulation is fp32 -> 1e-4 of the output scale (north_star allows 1e-3 for bf16/tf32 MLP
paths); measured errors are ~1e-5."""
TOL = 1e-4
import numpy as np
import pytest
import torch

from cpfn_b200 import fused

pytestmark = pytest.mark.gpu


def _chain(dims, rng, dev, relu_last=True):
    layers = []
    for i, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
        w = (rng.normal(size=(cout, cin)) * np.sqrt(2.0 / cin)).astype(np.float32)
        b = rng.normal(scale=0.1, size=cout).astype(np.float32)
        layers.append((w, b, relu_last or i < len(dims) - 2))
    return fused.PackedChain(layers, dev), layers


def _ref(x, layers):
    for w, b, relu in layers:
        x = x @ torch.from_numpy(w).to(x.device).t() + torch.from_numpy(b).to(x.device)
        if relu:
            x = torch.relu(x)
    return x


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-6))


@pytest.fixture(autouse=True)
def _fp32_reference():
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


@pytest.mark.parametrize("dims,tile", [([32, 128], 128), ([8, 64], 128), ([64, 64, 128], 128), ([128, 128, 128, 128, 128, 35], 128),
                                       ([100, 72, 40], 64), ([256, 256, 256], 64), ([1024, 256], 32), ([384, 256, 128], 64),
                                       ([260, 256, 512, 1024], 32)])
def test_dense_rows(cuda_dev, dims, tile):
    rng = np.random.default_rng(sum(dims))
    pc, layers = _chain(dims, rng, cuda_dev, relu_last=False)
    B, n = 3, 200          # 600 columns: ragged last tile
    x = torch.from_numpy(rng.normal(size=(B, n, dims[0])).astype(np.float32)).to(cuda_dev)
    out = torch.full((B, n, dims[-1]), float("nan"), device=cuda_dev)
    fused.run_chain(pc, B, n, out, dims[-1], tile_cols=tile, in_mode=fused.IN_DENSE, a_src=x, a_ch=dims[0], a_rows=n)
    ref = _ref(x, layers)
    assert torch.isfinite(out).all()
    assert _rel(out, ref) < TOL, _rel(out, ref)


@pytest.mark.parametrize("feat_ch,K,tile,S", [(0, 64, 128, 24), (64, 64, 128, 24), (128, 64, 64, 24), (128, 32, 64, 24),
                                              (256, 128, 32, 24),
                                              # points-as-M kernel: whole groups inside full tiles (pooled through shared
                                              # memory, plain stores) with 4 / 2 / 1 groups per tile ...
                                              (0, 32, 128, 24), (0, 128, 128, 24), (64, 64, 128, 26),
                                              # ... and ragged clouds (last tile partly empty: atomicMax path)
                                              (0, 32, 128, 25), (64, 64, 128, 25)])
def test_group_and_pool(cuda_dev, feat_ch, K, tile, S):
    rng = np.random.default_rng(feat_ch + K)
    B, N = 2, 500
    dims = [feat_ch + 3, 64, 128]
    pc, layers = _chain(dims, rng, cuda_dev)
    xyz = torch.from_numpy(rng.normal(size=(B, N, 3)).astype(np.float32)).to(cuda_dev)
    feats = torch.from_numpy(rng.normal(size=(B, N, feat_ch)).astype(np.float32)).to(cuda_dev) if feat_ch else None
    centers = torch.from_numpy(rng.normal(size=(B, S, 3)).astype(np.float32)).to(cuda_dev)
    idx = torch.from_numpy(rng.integers(0, N, size=(B, S, K)).astype(np.int32)).to(cuda_dev)
    out = torch.full((B, S, dims[-1]), float("nan"), device=cuda_dev)
    fused.run_chain(pc, B, S * K, out, dims[-1], tile_cols=tile, in_mode=fused.IN_GROUP, a_src=feats, a_ch=feat_ch,
                    a_rows=N, idx=idx, xyz=xyz, centers=centers, group_k=K, out_mode=fused.OUT_POOL, pool_g=K)
    li = idx.long()
    g_xyz = torch.gather(xyz, 1, li.reshape(B, S * K, 1).expand(-1, -1, 3)).reshape(B, S, K, 3) - centers[:, :, None, :]
    rows = g_xyz
    if feat_ch:
        g_f = torch.gather(feats, 1, li.reshape(B, S * K, 1).expand(-1, -1, feat_ch)).reshape(B, S, K, feat_ch)
        rows = torch.cat([g_f, g_xyz], dim=3)
    ref = _ref(rows, layers).max(dim=2)[0]
    assert _rel(out, ref) < TOL, _rel(out, ref)


def test_group_all_atomic_pool(cuda_dev):
    rng = np.random.default_rng(5)
    B, N = 3, 128
    dims = [259, 256, 512, 1024]
    pc, layers = _chain(dims, rng, cuda_dev)
    xyz = torch.from_numpy(rng.normal(size=(B, N, 3)).astype(np.float32)).to(cuda_dev)
    feats = torch.from_numpy(rng.normal(size=(B, N, 256)).astype(np.float32)).to(cuda_dev)
    idx = torch.arange(N, dtype=torch.int32, device=cuda_dev).repeat(B)
    out = torch.full((B, 1, 1024), float("nan"), device=cuda_dev)
    fused.run_chain(pc, B, N, out, 1024, tile_cols=32, in_mode=fused.IN_GROUP, a_src=feats, a_ch=256, a_rows=N, idx=idx,
                    xyz=xyz, centers=torch.zeros(B, 3, device=cuda_dev), group_k=N, out_mode=fused.OUT_POOL, pool_g=N)
    ref = _ref(torch.cat([feats, xyz], dim=2), layers).max(dim=1, keepdim=True)[0]
    assert _rel(out, ref) < TOL, _rel(out, ref)


@pytest.mark.parametrize("skip_ch,tile", [(0, 128), (128, 64)])
def test_interp_mask_and_channel_major_copy(cuda_dev, skip_ch, tile):
    rng = np.random.default_rng(11 + skip_ch)
    B, N, M, C2 = 2, 256, 40, 128
    dims = [skip_ch + C2, 128, 35]
    pc, layers = _chain(dims, rng, cuda_dev, relu_last=False)
    f2 = torch.from_numpy(rng.normal(size=(B, M, C2)).astype(np.float32)).to(cuda_dev)
    skip = torch.from_numpy(rng.normal(size=(B, N, skip_ch)).astype(np.float32)).to(cuda_dev) if skip_ch else None
    idx = torch.from_numpy(rng.integers(0, M, size=(B, N, 3)).astype(np.int32)).to(cuda_dev)
    w = torch.from_numpy(rng.uniform(size=(B, N, 3)).astype(np.float32)).to(cuda_dev)
    w = (w / w.sum(2, keepdim=True)).contiguous()
    keep = rng.integers(0, 2, size=(B, 128, N))
    mask = (torch.from_numpy(keep.astype(np.float32)) * 2).to(cuda_dev)
    # the kernel's form of the mask: one keep bit per (point, channel), word [b*N + n, c // 32], bit c % 32
    kb = keep.transpose(0, 2, 1).reshape(B * N, 4, 32).astype(np.uint64)
    bits = (kb << np.arange(32, dtype=np.uint64)).sum(axis=2).astype(np.uint32).view(np.int32)
    bits = torch.from_numpy(np.ascontiguousarray(bits)).to(cuda_dev)
    feat_cm = torch.full((B, 128, N), float("nan"), device=cuda_dev)
    out = torch.full((B, N, 35), float("nan"), device=cuda_dev)
    fused.run_chain(pc, B, N, out, 35, tile_cols=tile, in_mode=fused.IN_INTERP, a_src=skip, a_ch=skip_ch, a_rows=N,
                    idx=idx, b_src=f2, b_ch=C2, b_rows=M, nn_w=w, masks={0: (bits, 2.0)}, out_cm={0: feat_cm})
    g = torch.gather(f2, 1, idx.long().reshape(B, N * 3, 1).expand(-1, -1, C2)).reshape(B, N, 3, C2)
    interp = (g * w[..., None]).sum(2)
    rows = interp if skip is None else torch.cat([skip, interp], dim=2)
    h = torch.relu(rows @ torch.from_numpy(layers[0][0]).to(cuda_dev).t() + torch.from_numpy(layers[0][1]).to(cuda_dev))
    h = h * mask.permute(0, 2, 1)
    ref = h @ torch.from_numpy(layers[1][0]).to(cuda_dev).t() + torch.from_numpy(layers[1][1]).to(cuda_dev)
    assert _rel(feat_cm.permute(0, 2, 1), h) < TOL
    assert _rel(out, ref) < TOL, _rel(out, ref)


def test_per_cloud_bias_and_many_tiles(cuda_dev):
    """More tiles than CTAs (persistent loop, barrier phases wrap many times) + per-cloud bias."""
    rng = np.random.default_rng(3)
    B, n, dims = 16, 8192, [128, 128, 128]
    pc, layers = _chain(dims, rng, cuda_dev)
    x = torch.from_numpy(rng.normal(size=(B, n, 128)).astype(np.float32)).to(cuda_dev)
    bias0 = torch.from_numpy(rng.normal(size=(B, 128)).astype(np.float32)).to(cuda_dev)
    out = torch.empty(B, n, 128, device=cuda_dev)
    fused.run_chain(pc, B, n, out, 128, tile_cols=128, in_mode=fused.IN_DENSE, a_src=x, a_ch=128, a_rows=n,
                    biases=[bias0, None], bias_per_cloud=(0,))
    h = torch.relu(x @ torch.from_numpy(layers[0][0]).to(cuda_dev).t() + bias0[:, None, :])
    ref = torch.relu(h @ torch.from_numpy(layers[1][0]).to(cuda_dev).t() + torch.from_numpy(layers[1][1]).to(cuda_dev))
    assert _rel(out, ref) < TOL, _rel(out, ref)
    out2 = torch.empty_like(out)
    fused.run_chain(pc, B, n, out2, 128, tile_cols=128, in_mode=fused.IN_DENSE, a_src=x, a_ch=128, a_rows=n,
                    biases=[bias0, None], bias_per_cloud=(0,))
    assert torch.equal(out, out2)       # deterministic


def test_invalid_arguments(cuda_dev):
    rng = np.random.default_rng(0)
    pc, _ = _chain([16, 32], rng, cuda_dev)
    x = torch.zeros(1, 64, 16, device=cuda_dev)
    out = torch.zeros(1, 64, 32, device=cuda_dev)
    with pytest.raises(RuntimeError):
        fused.run_chain(pc, 1, 64, out, 32, tile_cols=48, a_src=x, a_ch=16, a_rows=64)          # bad tile
    with pytest.raises(RuntimeError):
        fused.run_chain(pc, 1, 64, out, 32, tile_cols=64, a_src=x, a_ch=12, a_rows=64)          # width mismatch


@pytest.mark.parametrize("dims,tile", [([512, 1024], 64), ([260, 256], 64), ([100, 300], 32)])
def test_single_layer_channel_split(cuda_dev, dims, tile):
    """split_cout: one CTA per (column tile, 128-channel chunk); rows and pooled outputs."""
    rng = np.random.default_rng(dims[1])
    pc, layers = _chain(dims, rng, cuda_dev)
    B, n = 4, 128
    x = torch.from_numpy(rng.normal(size=(B, n, dims[0])).astype(np.float32)).to(cuda_dev)
    out = torch.full((B, n, dims[1]), float("nan"), device=cuda_dev)
    fused.run_chain(pc, B, n, out, dims[1], tile_cols=tile, in_mode=fused.IN_DENSE, a_src=x, a_ch=dims[0], a_rows=n,
                    split_cout=True)
    ref = _ref(x, layers)
    assert _rel(out, ref) < TOL, _rel(out, ref)
    pooled = torch.full((B, 1, dims[1]), float("nan"), device=cuda_dev)
    fused.run_chain(pc, B, n, pooled, dims[1], tile_cols=tile, in_mode=fused.IN_DENSE, a_src=x, a_ch=dims[0], a_rows=n,
                    split_cout=True, out_mode=fused.OUT_POOL, pool_g=n)
    assert _rel(pooled, ref.max(dim=1, keepdim=True)[0]) < TOL


def test_layerwise_equals_fused_chain(cuda_dev):
    rng = np.random.default_rng(9)
    dims = [256, 256, 256]
    pc, layers = _chain(dims, rng, cuda_dev)
    B, n = 16, 128
    x = torch.from_numpy(rng.normal(size=(B, n, 256)).astype(np.float32)).to(cuda_dev)
    bias0 = torch.from_numpy(rng.normal(size=(B, 256)).astype(np.float32)).to(cuda_dev)
    a = torch.empty(B, n, 256, device=cuda_dev)
    fused.run_layerwise(pc, B, n, a, dict(in_mode=fused.IN_DENSE, a_src=x, a_ch=256, a_rows=n), bias0=bias0)
    h = torch.relu(x @ torch.from_numpy(layers[0][0]).to(cuda_dev).t() + bias0[:, None, :])
    ref = torch.relu(h @ torch.from_numpy(layers[1][0]).to(cuda_dev).t() + torch.from_numpy(layers[1][1]).to(cuda_dev))
    assert _rel(a, ref) < TOL
    w = torch.from_numpy(rng.normal(size=(40, 1000)).astype(np.float32)).to(cuda_dev)
    g = torch.from_numpy(rng.normal(size=(5, 1000)).astype(np.float32)).to(cuda_dev)
    bb = torch.from_numpy(rng.normal(size=40).astype(np.float32)).to(cuda_dev)
    o = torch.zeros(5, 128, device=cuda_dev)
    fused.linear_rows(g, w, bb, o)
    assert _rel(o[:, :40], g @ w.t() + bb) < 1e-5 and float(o[:, 40:].abs().max()) == 0.0


@pytest.mark.parametrize("feat_ch,K,S,cout", [(128, 64, 24, 256), (64, 32, 28, 128), (128, 128, 5, 200)])
def test_positions_as_epilogue_term(cuda_dev, feat_ch, K, S, cout):
    """GROUP mode with features on 128-column tiles: the three position columns of layer 0 enter as an fp32 term of
    its epilogue (cpfn_mlp_chain_t.xyz_w) and the pooled output leaves through shared memory -- same result as the
    reference composition and as the 64-column kernel with the positions inside the operand."""
    rng = np.random.default_rng(feat_ch + K + S)
    B, N = 2, 600
    dims = [feat_ch + 3, 128, 128, cout]
    pc, layers = _chain(dims, rng, cuda_dev)
    w0, b0, _ = layers[0]
    alt = fused.PackedChain([(np.ascontiguousarray(w0[:, :feat_ch]), b0, True)] + layers[1:], cuda_dev)
    rows = np.zeros((128, 4), dtype=np.float32)
    rows[:, :3], rows[:, 3] = w0[:, feat_ch:], b0
    xyz_w = torch.from_numpy(rows).to(cuda_dev)
    xyz = torch.from_numpy(rng.normal(size=(B, N, 3)).astype(np.float32)).to(cuda_dev)
    feats = torch.from_numpy(rng.normal(size=(B, N, feat_ch)).astype(np.float32)).to(cuda_dev)
    centers = torch.from_numpy(rng.normal(size=(B, S, 3)).astype(np.float32)).to(cuda_dev)
    idx = torch.from_numpy(rng.integers(0, N, size=(B, S, K)).astype(np.int32)).to(cuda_dev)
    out = torch.full((B, S, cout), float("nan"), device=cuda_dev)
    fused.run_chain(alt, B, S * K, out, cout, tile_cols=128, in_mode=fused.IN_GROUP, a_src=feats, a_ch=feat_ch, a_rows=N,
                    idx=idx, xyz=xyz, centers=centers, group_k=K, out_mode=fused.OUT_POOL, pool_g=K, xyz_w=xyz_w)
    old = torch.full((B, S, cout), float("nan"), device=cuda_dev)
    fused.run_chain(pc, B, S * K, old, cout, tile_cols=64, in_mode=fused.IN_GROUP, a_src=feats, a_ch=feat_ch, a_rows=N,
                    idx=idx, xyz=xyz, centers=centers, group_k=K, out_mode=fused.OUT_POOL, pool_g=K)
    li = idx.long()
    g_xyz = torch.gather(xyz, 1, li.reshape(B, S * K, 1).expand(-1, -1, 3)).reshape(B, S, K, 3) - centers[:, :, None, :]
    g_f = torch.gather(feats, 1, li.reshape(B, S * K, 1).expand(-1, -1, feat_ch)).reshape(B, S, K, feat_ch)
    ref = _ref(torch.cat([g_f, g_xyz], dim=3), layers).max(dim=2)[0]
    assert torch.isfinite(out).all()
    assert _rel(out, ref) < TOL, _rel(out, ref)
    assert _rel(out, old) < TOL, _rel(out, old)
    with pytest.raises(RuntimeError):                     # the term belongs to the 128-column kernel
        fused.run_chain(alt, B, S * K, out, cout, tile_cols=64, in_mode=fused.IN_GROUP, a_src=feats, a_ch=feat_ch,
                        a_rows=N, idx=idx, xyz=xyz, centers=centers, group_k=K, out_mode=fused.OUT_POOL, pool_g=K, xyz_w=xyz_w)


def test_first_layer_on_cuda_cores(cuda_dev):
    """GROUP mode on bare positions with the 3 -> 64 first layer evaluated in fp32 by the tile builder."""
    rng = np.random.default_rng(21)
    B, N, S, K = 2, 700, 40, 64
    dims = [3, 64, 64, 128]
    pc_full, layers = _chain(dims, rng, cuda_dev)
    pc = fused.PackedChain(layers[1:], cuda_dev)
    l0 = (torch.from_numpy(layers[0][0]).to(cuda_dev), torch.from_numpy(layers[0][1]).to(cuda_dev))
    xyz = torch.from_numpy(rng.normal(size=(B, N, 3)).astype(np.float32)).to(cuda_dev)
    centers = torch.from_numpy(rng.normal(size=(B, S, 3)).astype(np.float32)).to(cuda_dev)
    idx = torch.from_numpy(rng.integers(0, N, size=(B, S, K)).astype(np.int32)).to(cuda_dev)
    out = torch.full((B, S, 128), float("nan"), device=cuda_dev)
    fused.run_chain(pc, B, S * K, out, 128, tile_cols=128, in_mode=fused.IN_GROUP, a_src=None, a_ch=0, a_rows=N, idx=idx,
                    xyz=xyz, centers=centers, group_k=K, out_mode=fused.OUT_POOL, pool_g=K, l0=l0)
    g = torch.gather(xyz, 1, idx.long().reshape(B, S * K, 1).expand(-1, -1, 3)).reshape(B, S, K, 3) - centers[:, :, None, :]
    ref = _ref(g, layers).max(dim=2)[0]
    assert _rel(out, ref) < TOL, _rel(out, ref)
