"""Seam B1 of SURVEY 8b, live: the UNMODIFIED reference Python (PointNet2/pn2_network.py and its modules, staged
under baseline/_ref by oracle/build_ref.py) runs ``fast=True`` on this package's ``cuda_ops`` -- what
``cpfn_b200.dropin.install(level="ops")`` gives a user -- and is compared with

  * the same reference Python on the reference's own CUDA extension (oracle/_ref/ref_cuda_ops.so): every op is
    bit-exact, so the two forwards must agree bit for bit, dropout included (same torch generator state);
  * the fused whole-network path of this package (tcgen05 MLP chains, split-bf16): 2e-4 of the tensor scale,
    dropout included -- the fused path draws the mask torch's F.dropout draws.
"""
import sys

import numpy as np
import pytest
import torch

from cpfn_b200 import api, synth
from oracle import build_ref, ref_runtime
from tests.golden import cases

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.float().cpu().numpy(), b.float().cpu().numpy()
    return float(np.abs(a - b).max() / max(1e-6, np.abs(b).max()))


@pytest.fixture(scope="module")
def staged():
    if not ref_runtime.available():
        pytest.skip("baseline/_ref is not staged")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def _forward(net_module, sd, P, dev, seed):
    model = net_module.PointNet2(dim_input=3, dim_pos=3, output_sizes=[3, 4, 28]).to(dev)
    model.load_state_dict(sd, strict=True)                       # reference checkpoint layout, strict (training_SPFN.py:74)
    model.eval()
    with torch.no_grad():
        torch.manual_seed(seed)
        return [t.clone() for t in model(P, fast=True)]


@pytest.mark.parametrize("batch,n_points", [(2, 1024), (16, 8192)])
def test_reference_python_on_our_ops(cuda_dev, staged, batch, n_points):
    from cpfn_b200 import dropin
    eng = api.GlobalSPFN(output_sizes=[3, 4, 28], device=cuda_dev)
    sd = {k: torch.from_numpy(v) for k, v in synth.network_state(eng.model.state_dict(), seed=31).items()}
    eng.load_state_dict(sd, strict=True)
    P = torch.from_numpy(cases.network_input(batch=batch, n_points=n_points, seed=17)).to(cuda_dev)
    # dropin.install(level="ops") exactly as INTEGRATION.md describes it: reference tree on sys.path, only cuda_ops replaced
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in ("PointNet2", "SPFN", "Utils")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, ref_runtime.STAGE_DIR)
    try:
        dropin.install(level="ops")
        import PointNet2.pn2_network as ref_net
        import PointNet2.pointnet2_ops.modules.geometry_utils as ref_geo
        import cpfn_b200
        assert ref_geo.cuda_ops is cpfn_b200.cuda_ops and ref_net.__file__.startswith(ref_runtime.STAGE_DIR)
        on_ours = _forward(ref_net, sd, P, cuda_dev, seed=5)
    finally:
        sys.path.remove(ref_runtime.STAGE_DIR)
        for k in [k for k in sys.modules if k.split(".")[0] in ("PointNet2", "SPFN", "Utils")]:
            del sys.modules[k]
        sys.modules.update(saved)
    assert len(on_ours) == 5 and on_ours[2].shape == (batch, n_points, 28)
    # the fused path of this package with the same generator state
    torch.manual_seed(5)
    fused_out = eng.forward(P, dropout=True, fit=False)
    for i, k in enumerate(("X_raw", "T_raw", "W_raw")):
        assert _rel(fused_out[k], on_ours[i]) < 2e-4, k
    assert _rel(fused_out["l3_feats"], on_ours[3]) < 2e-4
    assert _rel(fused_out["output_feat"], on_ours[4]) < 2e-4
    # the very same dropout mask (zeros of the ReLU may differ where a pre-activation is within rounding of 0)
    assert float(((fused_out["output_feat"] == 0) != (on_ours[4] == 0)).float().mean()) < 1e-3
    # the same reference Python on the reference's own extension
    ref_ops = build_ref.load_module()
    if ref_ops is None:
        pytest.skip("oracle/_ref/ref_cuda_ops.so is not built")
    on_ref = _forward(ref_runtime.load_pointnet2(ref_ops), sd, P, cuda_dev, seed=5)
    for a, b in zip(on_ours, on_ref):
        assert torch.equal(a, b)
