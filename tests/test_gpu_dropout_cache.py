"""The dropout bit mask (cpfn_dropout_mask_bits), the staleness checks of the packed-weight caches and the lifetime of
the buffers baked into captured CUDA graphs.

The mask must be the one torch's own ``F.dropout`` draws for the same generator state (the reference calls
``F.dropout(x, p=0.5)`` with training=True on a [B,128,N] tensor, PointNet2/pn2_network.py:63): same Philox stream,
same element -> (thread, iteration, component) assignment, same generator advance."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from cpfn_b200 import api, fused, synth
from cpfn_b200.pn2_network import PointNet2
from tests.golden import cases

pytestmark = pytest.mark.gpu


def _unpack(bits, B, N, C):
    words = (C + 31) // 32
    w = bits.cpu().numpy().view(np.uint32).reshape(B, N, words)
    keep = ((w[..., None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(B, N, words * 32)[:, :, :C]
    return keep.transpose(0, 2, 1).astype(bool)


@pytest.mark.parametrize("B,C,N", [(2, 128, 1024), (16, 128, 8192), (3, 128, 1001), (1, 128, 20000), (2, 96, 515)])
def test_mask_bits_equal_torch_dropout(cuda_dev, B, C, N):
    torch.manual_seed(1234)
    torch.rand(7, device=cuda_dev)                                    # a non-zero generator offset
    want = F.dropout(torch.ones(B, C, N, device=cuda_dev), p=0.5) != 0
    after_torch = torch.rand(5, device=cuda_dev)
    torch.manual_seed(1234)
    torch.rand(7, device=cuda_dev)
    bits, scale = fused.dropout_bits(B, C, N, cuda_dev, p=0.5)
    after_ours = torch.rand(5, device=cuda_dev)
    got = _unpack(bits, B, N, C)
    assert scale == 2.0
    assert np.array_equal(got, want.cpu().numpy())
    assert torch.equal(after_torch, after_ours)                       # the generator moved on exactly as torch's call does


def test_mask_is_fresh_and_seeded(cuda_dev):
    torch.manual_seed(5)
    a = fused.dropout_bits(2, 128, 512, cuda_dev)[0].clone()
    b = fused.dropout_bits(2, 128, 512, cuda_dev)[0].clone()
    torch.manual_seed(5)
    c = fused.dropout_bits(2, 128, 512, cuda_dev)[0].clone()
    assert not torch.equal(a, b) and torch.equal(a, c)


@pytest.fixture(scope="module")
def weights(cuda_dev):
    tmpl = PointNet2(output_sizes=[3, 4, 28]).state_dict()
    return [{k: torch.from_numpy(v) for k, v in synth.network_state(tmpl, seed=s).items()} for s in (21, 22)]


def test_packed_weight_cache_follows_the_parameters(cuda_dev, weights):
    """train -> eval -> train -> eval (the reference's loop validates under .eval() + no_grad between epochs,
    training_SPFN.py): the fused eval path must see the current parameters and BatchNorm statistics each time."""
    torch.backends.cudnn.allow_tf32 = False
    P = torch.from_numpy(cases.network_input(batch=2, n_points=1024)).to(cuda_dev)
    model = PointNet2(output_sizes=[3, 4, 28]).to(cuda_dev)
    fresh = PointNet2(output_sizes=[3, 4, 28]).to(cuda_dev)

    def eval_out(m):
        m.eval()
        with torch.no_grad():
            return m.sa1(P.transpose(1, 2).contiguous(), None)[1].clone(), m(P)[3].clone()

    def same_as_fresh_copy():
        fresh.load_state_dict(model.state_dict(), strict=True)
        fused.invalidate(fresh)
        a, b = eval_out(model), eval_out(fresh)
        return torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])

    model.load_state_dict(weights[0], strict=True)
    first = eval_out(model)
    assert same_as_fresh_copy()
    model.load_state_dict(weights[1], strict=True)                    # load_state_dict on the drop-in module itself
    second = eval_out(model)
    assert not torch.equal(first[1], second[1]) and same_as_fresh_copy()
    model.train()                                                     # one optimiser step + BatchNorm statistics update
    opt = torch.optim.SGD(model.parameters(), lr=0.05)
    loss = sum(r.float().pow(2).mean() for r in model(P)[:3])
    loss.backward()
    opt.step()
    third = eval_out(model)
    assert not torch.equal(second[1], third[1]) and same_as_fresh_copy()
    with torch.no_grad():
        model.bn1.running_var.mul_(1.5)                               # a buffer edited in place
        model.fc2[2].bias.add_(0.25)
    model.eval()
    with torch.no_grad():
        h = model(P)[2]
    fresh.load_state_dict(model.state_dict(), strict=True)
    fresh.eval()
    with torch.no_grad():
        torch.manual_seed(1)
        a = model(P)[2]
        torch.manual_seed(1)
        b = fresh(P)[2]
    assert torch.equal(a, b) and h.shape == a.shape


def test_engine_graphs_follow_the_weights(cuda_dev, weights):
    eng = api.GlobalSPFN(output_sizes=[3, 4, 28], device=cuda_dev)
    eng.load_state_dict(weights[0])
    P = torch.from_numpy(cases.network_input()).to(cuda_dev)
    a = eng.forward_graphed(P, dropout=False)["W_raw"].clone()
    eng.model.load_state_dict(weights[1])                             # behind the engine's back
    b = eng.forward_graphed(P, dropout=False)["W_raw"].clone()
    want = eng.forward(P, dropout=False)["W_raw"]
    assert not torch.equal(a, b) and torch.equal(b, want)


def test_graph_replay_after_capturing_another_shape(cuda_dev, weights):
    """Buffers baked into a captured graph (per-cloud bias rows, fitter scratch, ball-query grids, the dropout
    bits) must live as long as the graph: capture shape A, capture B and C (larger, then smaller), exercise the
    allocator, replay A."""
    eng = api.GlobalSPFN(output_sizes=[3, 4, 28], device=cuda_dev)
    eng.load_state_dict(weights[0])
    PA = torch.from_numpy(cases.network_input(batch=2, n_points=2048, seed=5)).to(cuda_dev)
    want = {k: v.clone() for k, v in eng.forward(PA, dropout=False)["parameters"].items()}
    heads = eng.forward(PA, dropout=False)["W_raw"].clone()
    got = eng.forward_graphed(PA, dropout=False)
    assert torch.equal(got["W_raw"], heads)
    for shape in ((5, 4096), (1, 1024), (3, 2048)):
        Pb = torch.from_numpy(cases.network_input(batch=shape[0], n_points=shape[1], seed=9)).to(cuda_dev)
        eng.forward_graphed(Pb, dropout=True)
        eng.forward(Pb, dropout=True)
    junk = [torch.full((1 << 20,), float("nan"), device=cuda_dev) for _ in range(64)]   # recycle freed blocks
    del junk
    torch.cuda.synchronize()
    again = eng.forward_graphed(PA, dropout=False)
    assert torch.equal(again["W_raw"], heads)
    for k, v in want.items():
        assert torch.equal(again["parameters"][k], v), k
    m1 = eng.forward_graphed(PA, dropout=True)["output_feat"].clone()
    m2 = eng.forward_graphed(PA, dropout=True)["output_feat"].clone()
    frac = float((m1 == 0).float().mean())
    assert not torch.equal(m1, m2) and 0.45 < frac < 0.8
