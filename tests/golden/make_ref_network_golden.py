"""Generate tests/golden/ref_network.npz by running the UNMODIFIED reference network
(PointNet2/pn2_network.py and its pointnet2_ops/modules) on CPU in the dev container.

    python tests/golden/make_ref_network_golden.py

The reference modules call ``cuda_ops`` (a CUDA-only extension) for FPS / ball query /
three_nn / three_weighted_sum.  No GPU exists here, so a stand-in module exposing the
same nine functions on CPU tensors is registered as ``PointNet2.pointnet2_ops.cuda_ops``
before the import; it forwards to oracle/cpfn_oracle.c, which is itself pinned bit-exactly
to the reference kernels (tests/golden/ref_cuda_ops.npz, made on a B200 from the
reference .so).  Everything else -- module wiring, grouping order, re-centring, the
conv/BN/ReLU/max chain, interpolation weights, heads -- is the reference's own code.
Dropout (pn2_network.py:63) draws from torch's CPU generator; the same mask is
reproduced with the same seed and stored, and consistency is asserted.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
from oracle import index_ops  # noqa: E402
from tests.golden import cases  # noqa: E402


def install_cuda_ops_stand_in():
    m = types.ModuleType("PointNet2.pointnet2_ops.cuda_ops")
    t = torch.from_numpy
    m.farthest_point_sampling = lambda p, n: t(index_ops.farthest_point_sampling(p.numpy(), n))
    m.ball_query = lambda q, x, r, k: t(index_ops.ball_query(q.numpy(), x.numpy(), r, k))
    m.three_nn = lambda u, k: [t(a) for a in index_ops.three_nn(u.numpy(), k.numpy())]
    m.three_weighted_sum = lambda p, i, w: t(index_ops.three_weighted_sum(p.numpy(), i.numpy(), w.numpy()))
    import PointNet2.pointnet2_ops as pkg  # noqa: F401  (the package's __init__ is empty)
    sys.modules["PointNet2.pointnet2_ops.cuda_ops"] = m
    pkg.cuda_ops = m


def main():
    torch.solve = lambda B, A: (torch.linalg.solve(A, B), None)   # SPFN import chain (torch 2.11)
    install_cuda_ops_stand_in()
    from PointNet2.pn2_network import PointNet2
    torch.set_num_threads(4)
    torch.backends.mkldnn.enabled = False        # plain fp32 convolutions
    model = PointNet2(dim_input=3, dim_pos=3, output_sizes=[3, 4, 28]).eval()
    state = cases.network_state(model.state_dict())
    model.load_state_dict({k: torch.from_numpy(v) for k, v in state.items()}, strict=True)
    P = cases.network_input()
    cap = {}
    model.bn1.register_forward_hook(lambda mod, i, o: cap.__setitem__("bn1", o.detach().clone()))
    model.sa1.register_forward_hook(lambda mod, i, o: cap.__setitem__("l1", o[1].detach().clone()))
    model.sa2.register_forward_hook(lambda mod, i, o: cap.__setitem__("l2", o[1].detach().clone()))
    model.sfp3.register_forward_hook(lambda mod, i, o: cap.__setitem__("l6", o.detach().clone()))
    with torch.no_grad():
        torch.manual_seed(99)
        res = model(torch.from_numpy(P))
        torch.manual_seed(99)
        mask = torch.nn.functional.dropout(torch.ones_like(cap["bn1"]), p=0.5)
    pre = torch.relu(cap["bn1"])
    assert torch.equal(pre * mask, res[-1]), "dropout mask reconstruction failed"
    out = {
        "head0": res[0].numpy(), "head1": res[1].numpy(), "head2": res[2].numpy(),
        "l3_feats": res[3].numpy()[:, :, 0],
        "mask_bits": np.packbits((mask.numpy() > 0).astype(np.uint8)),
        "feat_pre_dropout_s": pre.numpy()[:, ::8, ::4],
        "l1_feats_s": cap["l1"].numpy()[:, ::8, ::4], "l2_feats_s": cap["l2"].numpy()[:, ::8, ::2],
        "l6_feats_s": cap["l6"].numpy()[:, ::8, ::4],
    }
    path = os.path.join(ROOT, "tests", "golden", "ref_network.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
