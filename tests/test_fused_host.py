"""Host-side preparation of the fused chains (cpfn_b200/fused.py) that needs no GPU: BatchNorm folding, the packed
weight image, and the split of a set-abstraction chain's first layer into its feature columns (tensor-core operand)
and its three position columns (epilogue term, cpfn_mlp_chain_t.xyz_w)."""
import numpy as np
import torch

from cpfn_b200 import fused
from cpfn_b200.pointnet2_ops.modules.pointset_abstraction import PointsetAbstraction


def _randomise(module, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for p in module.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * 0.2)
        for name, b in module.named_buffers():
            if name.endswith("running_var"):
                b.copy_(torch.rand(b.shape, generator=g) + 0.5)
            elif name.endswith("running_mean"):
                b.copy_(torch.randn(b.shape, generator=g) * 0.1)


def test_sa_chain_splits_position_columns(built_lib):
    sa2 = PointsetAbstraction(num_points=128, dim_pos=3, dim_feats=128, radius_list=[0.4], num_samples_list=[64],
                              mlp_list=[[128, 128, 256]], group_all=False).eval()
    _randomise(sa2, 3)
    pc = fused._sa_chain(sa2, torch.device("cpu"))
    assert pc.dims == [(131, 128, True), (128, 128, True), (128, 256, True)] and pc.l0 is None
    alt = pc.alt
    assert alt is not None and alt.dims == [(128, 128, True), (128, 128, True), (128, 256, True)]
    w0, b0 = fused.fold_bn(sa2.conv_blocks[0][0].weight, sa2.conv_blocks[0][0].bias, sa2.bn_blocks[0][0])
    rows = alt.xyz_w.numpy()
    assert rows.shape == (128, 4)
    np.testing.assert_array_equal(rows[:, :3], w0[:, 128:131])          # reference column order: [features, xyz]
    np.testing.assert_array_equal(rows[:, 3], b0)
    # the operand part is the packed image of the feature columns alone: one K atom less than the full layer
    assert alt.weights.numel() == pc.weights.numel() - 2 * 16384
    np.testing.assert_array_equal(alt.weights.numpy()[: 4 * 16384],
                                  fused.pack_weights(np.ascontiguousarray(w0[:, :128])))
    # the cache entry follows the parameters: an in-place update rebuilds both chains
    with torch.no_grad():
        sa2.conv_blocks[0][0].weight.mul_(2.0)
    pc2 = fused._sa_chain(sa2, torch.device("cpu"))
    assert pc2 is not pc
    np.testing.assert_allclose(pc2.alt.xyz_w.numpy()[:, :3], 2.0 * rows[:, :3], rtol=1e-6)


def test_bare_position_layer_stays_on_the_cuda_cores(built_lib):
    sa1 = PointsetAbstraction(num_points=512, dim_pos=3, dim_feats=0, radius_list=[0.2], num_samples_list=[64],
                              mlp_list=[[64, 64, 128]], group_all=False).eval()
    _randomise(sa1, 4)
    pc = fused._sa_chain(sa1, torch.device("cpu"))
    assert pc.alt is None and pc.l0 is not None                        # 3 -> 64 first layer: fp32 in the tile builder
    assert pc.dims == [(64, 64, True), (64, 128, True)]
    assert tuple(pc.l0[0].shape) == (64, 3) and tuple(pc.l0[1].shape) == (64,)
