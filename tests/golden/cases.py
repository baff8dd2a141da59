"""Seeded inputs shared by the golden-vector generators and the parity tests.

Inputs are regenerated from seeds (cpfn_b200.synth, numpy default_rng), only the
reference OUTPUTS are stored in the .npz fixtures.
"""
import numpy as np

from cpfn_b200 import synth


def fps_cases():
    """name -> (xyz [B,N,3], nsamples)."""
    P_shape = synth.shape_batch(2, 8192, seed=11)[0]
    return {
        "uniform_8192_512": (synth.uniform_cloud(2, 8192, seed=1), 512),
        "shape_8192_512": (P_shape, 512),
        "lattice_1000_300": (synth.lattice_cloud(3, 1000, seed=2), 300),
        "lattice_259_100": (synth.lattice_cloud(2, 259, seed=3), 100),
        "uniform_512_128": (synth.uniform_cloud(4, 512, seed=4), 128),
        "lattice_40_40": (synth.lattice_cloud(2, 40, seed=5, pitch=4), 40),
        "uniform_12000_64": (synth.uniform_cloud(1, 12000, seed=6), 64),
        "uniform_20000_48": (synth.uniform_cloud(1, 20000, seed=7), 48),
    }


def _take(xyz, idx):
    return np.take_along_axis(xyz, idx[..., None].astype(np.int64), 1)


def ball_cases(fps):
    """name -> (new_xyz, xyz, radius, nsample); `fps(xyz, m)` supplies centroids."""
    u = synth.uniform_cloud(2, 8192, seed=1)
    s = synth.shape_batch(2, 8192, seed=11)[0]
    lat = synth.lattice_cloud(2, 1000, seed=2)
    c1 = _take(s, fps(s, 512))
    cases = {
        "sa1_shape": (c1, s, 0.2, 64),
        "sa2_shape": (_take(c1, fps(c1, 128)), c1, 0.4, 64),
        "sa1_uniform": (_take(u, fps(u, 512))[:, :128], u, 0.2, 64),
        "lattice_ties": (lat[:, :200], lat, 0.25, 32),
        "tiny_radius": (u[:, :64] + np.float32(0.5), u, 1e-3, 16),
        "sparse_hits": (u[:, :64], u[:, :300], 0.2, 48),
    }
    return cases


def three_nn_cases(fps):
    s = synth.shape_batch(1, 4096, seed=12)[0]
    lat = synth.lattice_cloud(1, 1000, seed=8, pitch=6)
    return {
        "fp3_like": (s, _take(s, fps(s, 512))),
        "lattice_ties": (lat, lat[:, :100]),
        "two_known": (s[:, :50], s[:, :2]),
    }


def interp_inputs():
    rng = np.random.default_rng(21)
    B, C, M, n = 2, 16, 128, 1024
    pts = rng.normal(size=(B, C, M)).astype(np.float32)
    idx = rng.integers(0, M, size=(B, n, 3)).astype(np.int32)
    w = rng.random(size=(B, n, 3)).astype(np.float32)
    w /= w.sum(2, keepdims=True)
    g = rng.normal(size=(B, C, n)).astype(np.float32)
    gidx = rng.integers(0, M, size=(B, 32, 8)).astype(np.int32)
    gg = rng.normal(size=(B, C, 32, 8)).astype(np.float32)
    return dict(pts=pts, idx=idx, w=w, g=g, gidx=gidx, gg=gg, M=M)
