"""Parity at the configurations bench.py times (BASELINE.json configs[1] and configs[2]'s network):

  * GlobalSPFN forward at B = 16 x 8192 points, heads [3, 4, 28]: head outputs and ``output_feat`` against the CPU
    oracle (oracle/network.py), sampled centroids bit-exact, and the same forward WITH the always-on dropout
    against the oracle fed with the very mask the GPU drew;
  * the primitive parameters that come out of network -> softmax / normalise -> four fitters, against the numpy
    oracle of the fitters (oracle/fitters.py) fed with the same memberships (1e-5, the north star's fp32 TLS bar)
    and against the all-oracle pipeline (1e-3, its bar for the bf16 / tf32 MLP path);
  * the PatchSelection network (heads [2], evaluation_PatchSelection.py:57-88) through the engine and through the
    reference-named module.
"""
import numpy as np
import pytest
import torch

from cpfn_b200 import api, fused, synth
from oracle import fitters as ofit
from oracle import network as onet
from tests.golden import cases

pytestmark = pytest.mark.gpu
SIGN_FREE = ("plane_normal", "cylinder_axis")
FLIPS_WITH = {"plane_center": "plane_normal"}


def _rel(a, b):
    return float(np.abs(a - b).max() / max(1e-6, np.abs(b).max()))


def _unpack_bits(bits, B, N, C=128):
    """int32 [B*N, C/32] keep bits -> float mask [B, C, N] with values 0 / 2 (the oracle's dropout_mask)."""
    w = bits.cpu().numpy().view(np.uint32).reshape(B, N, C // 32)
    keep = (w[..., None] >> np.arange(32, dtype=np.uint32)) & 1
    return (2.0 * keep.reshape(B, N, C).transpose(0, 2, 1)).astype(np.float32)


@pytest.fixture(scope="module")
def bench_forward(cuda_dev):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    eng = api.GlobalSPFN(output_sizes=[3, 4, 28], device=cuda_dev)
    sd = {k: torch.from_numpy(v) for k, v in synth.network_state(eng.model.state_dict(), seed=1234).items()}
    eng.load_state_dict(sd, strict=True)
    P = synth.shape_batch(16, 8192, seed=1234, k_slots=28)[0]            # bench.py's first batch
    out = eng.forward(torch.from_numpy(P).to(cuda_dev), dropout=False)
    ref = onet.pointnet2_forward(sd, P, 3)
    return eng, sd, P, out, ref


def test_network_at_bench_config(bench_forward):
    eng, sd, P, out, ref = bench_forward
    l1 = np.take_along_axis(P, ref["sa1_fps"].astype(np.int64)[:, :, None], axis=1)
    assert np.array_equal(out["l1_pos"].permute(0, 2, 1).cpu().numpy(), l1)          # FPS picks, bit-exact
    l2 = np.take_along_axis(l1, ref["sa2_fps"].astype(np.int64)[:, :, None], axis=1)
    assert np.array_equal(out["l2_pos"].permute(0, 2, 1).cpu().numpy(), l2)
    assert _rel(out["l3_feats"].cpu().numpy(), ref["l3_feats"]) < 2e-4
    assert _rel(out["output_feat"].cpu().numpy(), ref["feat_pre_dropout"]) < 2e-4
    for i, k in enumerate(("X_raw", "T_raw", "W_raw")):
        assert _rel(out[k].cpu().numpy(), ref["heads"][i]) < 2e-4, k
    Xn, _, Wn = onet.spfn_postprocess(ref["heads"])
    assert np.abs(out["W"].cpu().numpy() - Wn).max() < 1e-4
    assert np.abs(out["X"].cpu().numpy() - Xn).max() < 1e-3


def test_network_with_dropout_at_bench_config(bench_forward, cuda_dev):
    """The always-on dropout (pn2_network.py:63): heads against the oracle given the SAME mask."""
    eng, sd, P, _, _ = bench_forward
    B, N = P.shape[:2]
    torch.manual_seed(77)
    out = eng.forward(torch.from_numpy(P).to(cuda_dev), dropout=True, fit=False)
    torch.manual_seed(77)
    bits, scale = fused.dropout_bits(B, 128, N, cuda_dev, p=0.5)          # the same draw
    mask = _unpack_bits(bits, B, N)
    assert scale == 2.0 and 0.49 < (mask != 0).mean() < 0.51
    ref = onet.pointnet2_forward(sd, P, 3, dropout_mask=mask)
    assert _rel(out["output_feat"].cpu().numpy(), ref["output_feat"]) < 2e-4
    for i, k in enumerate(("X_raw", "T_raw", "W_raw")):
        assert _rel(out[k].cpu().numpy(), ref["heads"][i]) < 2e-4, k


def _param_err(got, ref):
    """{key: [B,K] error relative to max(1, |ref|_inf)} with the sign conventions of SURVEY A.7."""
    err = {}
    for key, val in got.items():
        a, b = val.astype(np.float64).copy(), ref[key].astype(np.float64)
        if key in SIGN_FREE:
            a = a * np.sign(np.sum(a * b, axis=-1, keepdims=True))
        if key in FLIPS_WITH:
            o = FLIPS_WITH[key]
            a = a * np.sign(np.sum(got[o].astype(np.float64) * ref[o], axis=-1))
        d = np.abs(a - b)
        err[key] = (d.max(axis=-1) if d.ndim == 3 else d) / max(1.0, float(np.abs(b).max()))
    return err


WELL_POSED = ("plane_normal", "plane_center", "sphere_center", "sphere_radius_squared")


def test_fitted_parameters_at_bench_config(bench_forward):
    """Network -> softmax -> fit, end to end, all 16 x 28 slots, all ten keys.
      * against the all-oracle pipeline (oracle network + oracle fitters): 1e-3, the north star's bar for results
        downstream of the bf16 / tf32 MLP path;
      * against the fitters' oracle on the SAME memberships / normals: 1e-5 (its fp32 TLS bar) for the plane and
        sphere keys outright.  The memberships of a randomly initialised network are near-uniform, which makes the
        cylinder and cone fits on them ill conditioned (axis = smallest eigenvector of a nearly isotropic normal
        scatter): there the yardstick is the reference arithmetic itself -- the float32 oracle's distance from the
        same algebra in float64 -- and the GPU result must be no further from that float64 answer than twice what
        the reference's own float32 arithmetic is (or 1e-5)."""
    eng, sd, P, out, ref = bench_forward
    got = {k: v.cpu().numpy() for k, v in out["parameters"].items()}
    Wg, Xg = out["W"].cpu().numpy(), out["X"].cpu().numpy()
    Xn, _, Wn = onet.spfn_postprocess(ref["heads"])
    all_oracle = ofit.compute_parameters(P, Wn, Xn)            # oracle network + oracle fitters
    e_all = _param_err(got, all_oracle)
    for key in got:
        assert e_all[key].max() < 1e-3, (key, float(e_all[key].max()))
    same_input = ofit.compute_parameters(P, Wg, Xg)            # oracle fitters on the GPU's memberships / normals
    exact = ofit.compute_parameters_f64(P, Wg, Xg)
    e_same, e_gpu, e_ref = _param_err(got, same_input), _param_err(got, exact), _param_err(same_input, exact)
    for key in got:
        if key in WELL_POSED:
            assert e_same[key].max() < 1e-5, (key, float(e_same[key].max()))
        assert e_gpu[key].max() <= max(1e-5, 2 * e_ref[key].max()), (key, float(e_gpu[key].max()), float(e_ref[key].max()))
        assert np.median(e_gpu[key]) <= max(1e-6, 2 * np.median(e_ref[key])), key
    packed = out["parameters_packed"].cpu().numpy()
    assert packed.size == 22 * 16 * 28 and np.isfinite(packed).all()


@pytest.mark.parametrize("n_points,batch", [(8192, 1), (8192, 4)])
def test_patch_selection_network(cuda_dev, n_points, batch):
    """PointNet2(output_sizes=[2]) as evaluation_PatchSelection.py:57-88 runs it (low-res cloud of 8192 points,
    argmax of the two logits per point): engine and reference-named module against the oracle."""
    from cpfn_b200.pn2_network import PointNet2
    torch.backends.cudnn.allow_tf32 = False
    eng = api.GlobalSPFN(output_sizes=[2], device=cuda_dev)
    sd = {k: torch.from_numpy(v) for k, v in synth.network_state(eng.model.state_dict(), seed=5).items()}
    eng.load_state_dict(sd, strict=True)
    assert sum(v.numel() for k, v in sd.items() if "num_batches" not in k and "running" not in k) == 1402050
    P = cases.network_input(batch=batch, n_points=n_points, seed=90 + batch)
    ref = onet.pointnet2_forward(sd, P, 1)
    out = eng.forward(torch.from_numpy(P).to(cuda_dev), dropout=False, fit=False)
    assert out["X_raw"].shape == (batch, n_points, 2)
    assert _rel(out["X_raw"].cpu().numpy(), ref["heads"][0]) < 2e-4
    assert _rel(out["output_feat"].cpu().numpy(), ref["feat_pre_dropout"]) < 2e-4
    logits = ref["heads"][0]
    margin = np.abs(logits[..., 0] - logits[..., 1]) > 1e-3 * np.abs(logits).max()
    pred = out["X_raw"].argmax(dim=2).cpu().numpy()
    assert np.array_equal(pred[margin], logits.argmax(2)[margin])          # the labels the script saves
    # the drop-in module (same signature / output list as the reference's PointNet2.forward), dropout on
    model = PointNet2(dim_input=3, dim_pos=3, output_sizes=[2]).to(cuda_dev).eval()
    model.load_state_dict(sd, strict=True)
    with torch.no_grad():
        torch.manual_seed(3)
        res = model(torch.from_numpy(P).to(cuda_dev))
        torch.manual_seed(3)
        bits, _ = fused.dropout_bits(batch, 128, n_points, cuda_dev)
    assert len(res) == 3 and res[0].shape == (batch, n_points, 2) and res[2].shape == (batch, 128, n_points)
    ref_d = onet.pointnet2_forward(sd, P, 1, dropout_mask=_unpack_bits(bits, batch, n_points))
    assert _rel(res[0].cpu().numpy(), ref_d["heads"][0]) < 2e-4
    assert _rel(res[1].cpu().numpy(), ref["l3_feats"]) < 2e-4
