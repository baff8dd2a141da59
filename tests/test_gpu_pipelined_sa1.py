"""The pipelined first set-abstraction layer (fused.sa1_pipelined): furthest point sampling split over several
launches, ball query on centroid sub-ranges and the MLP chain on column windows must reproduce the one-launch
results bit for bit -- indices, centroids, pooled features and the whole forward."""
import ctypes
import os

import numpy as np
import pytest
import torch

from cpfn_b200 import _lib, api, cuda_ops, fused, synth
from tests.golden import cases

pytestmark = pytest.mark.gpu


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


@pytest.mark.parametrize("B,N,m,cuts", [(16, 8192, 512, (0, 128, 256, 384, 512)), (3, 4096, 300, (0, 1, 2, 150, 299, 300)),
                                        (5, 2500, 64, (0, 64)), (2, 16384, 100, (0, 37, 100))])
def test_fps_rounds_equal_one_launch(cuda_dev, B, N, m, cuts):
    L = _lib.lib()
    P = torch.from_numpy(synth.shape_batch(B, N, seed=7)[0] if N != 2500 else synth.lattice_cloud(B, N, seed=3)).to(cuda_dev)
    if not L.cpfn_fps_rounds_supported(B, N):
        pytest.skip("outside the cluster kernel's domain")
    want_idx, want_xyz = cuda_ops.farthest_point_sampling(P, m, return_centroids=True)
    idx = torch.full((B, m), -1, dtype=torch.int32, device=cuda_dev)
    xyz = torch.full((B, m, 3), float("nan"), device=cuda_dev)
    state = torch.full((B, N), float("nan"), device=cuda_dev)
    for j0, j1 in zip(cuts[:-1], cuts[1:]):
        _lib.check(L.cpfn_furthest_point_sampling_rounds(P.data_ptr(), B, N, m, j0, j1, idx.data_ptr(), xyz.data_ptr(),
                                                         state.data_ptr(), state.numel() * 4, 120 * 1024 if j0 else 0,
                                                         _stream(cuda_dev)), "rounds")
    assert torch.equal(idx, want_idx) and torch.equal(xyz, want_xyz)


def test_ball_query_range_and_chain_window(cuda_dev):
    from cpfn_b200.pn2_network import PointNet2
    model = PointNet2(output_sizes=[3, 4, 28]).to(cuda_dev).eval()
    model.load_state_dict({k: torch.from_numpy(v) for k, v in synth.network_state(model.state_dict(), seed=4).items()})
    P = torch.from_numpy(synth.shape_batch(4, 8192, seed=21)[0]).to(cuda_dev)
    with torch.no_grad():
        new_xyz, want = fused.sa_forward_pm(model.sa1, P, None)
        os.environ["CPFN_SA1_CHUNKS"] = "4"
        got = fused.sa1_pipelined(model.sa1, P, torch.cuda.Stream(device=cuda_dev), 4)
        assert got is not None
        torch.cuda.current_stream(cuda_dev).wait_event(got[2])
        assert torch.equal(got[0], new_xyz) and torch.equal(got[1], want)
        got8 = fused.sa1_pipelined(model.sa1, P, torch.cuda.Stream(device=cuda_dev), 8)
        torch.cuda.current_stream(cuda_dev).wait_event(got8[2])
        assert torch.equal(got8[1], want)
        assert fused.sa1_pipelined(model.sa1, P, torch.cuda.Stream(device=cuda_dev), 3) is None      # 512 % 3


def test_forward_is_identical_with_and_without_the_pipeline(cuda_dev):
    eng = api.GlobalSPFN(output_sizes=[3, 4, 28], device=cuda_dev)
    eng.load_state_dict({k: torch.from_numpy(v) for k, v in synth.network_state(eng.model.state_dict(), seed=9).items()})
    P = torch.from_numpy(synth.shape_batch(16, 8192, seed=1234)[0]).to(cuda_dev)
    outs = {}
    try:
        for chunks in ("1", "4", "2"):
            os.environ["CPFN_SA1_CHUNKS"] = chunks
            torch.manual_seed(3)
            o = eng.forward(P, dropout=True)
            outs[chunks] = {k: o[k].clone() for k in ("W_raw", "X_raw", "T_raw", "output_feat", "l3_feats")}
            outs[chunks].update({k: v.clone() for k, v in o["parameters"].items()})
            eng._graphs.clear()
            eng.forward_graphed(P, dropout=True)                       # capture (its warm-up runs draw masks too)
            torch.manual_seed(3)
            g = eng.forward_graphed(P, dropout=True)                   # the same through a CUDA graph (three streams)
            for k in ("W_raw", "output_feat"):
                assert torch.equal(g[k], outs[chunks][k]), (chunks, k)
    finally:
        os.environ.pop("CPFN_SA1_CHUNKS", None)
    for chunks in ("4", "2"):
        for k, v in outs["1"].items():
            assert torch.equal(outs[chunks][k], v), (chunks, k)
