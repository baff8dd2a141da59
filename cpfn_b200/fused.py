"""Host side of the fused tcgen05 MLP-chain kernel (csrc/mlp_chain.cu, C ABI
``cpfn_mlp_chain``): BatchNorm folding and weight packing at load time, and the
inference forward of PointsetAbstraction / PointsetFeaturePropagation / the whole
PointNet2 on top of it.

Layout: between the fused kernels features are POINT-major ([B, N, C], a point's
channels contiguous) so that a neighbourhood gather reads whole 512-byte rows; the
reference's channel-major [B, C, N] tensors are produced only where the reference API
returns them (``l3_feats``, ``output_feat``, the module-level SA / FP forwards).
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib, cuda_ops

MAX_LAYERS = 6
IN_DENSE, IN_GROUP, IN_INTERP = 0, 1, 2
OUT_ROWS, OUT_POOL = 0, 1
TIMED_OPS = ("mlp_chain",)


class _Layer(ctypes.Structure):
    _fields_ = [("cin", ctypes.c_int32), ("cout", ctypes.c_int32), ("relu", ctypes.c_int32),
                ("bias_per_cloud", ctypes.c_int32), ("bias", ctypes.c_void_p), ("mask_bits", ctypes.c_void_p),
                ("mask_scale", ctypes.c_float), ("out_cm", ctypes.c_void_p)]


class _Chain(ctypes.Structure):
    _fields_ = [("n_layers", ctypes.c_int32), ("layers", _Layer * MAX_LAYERS),
                ("weights", ctypes.c_void_p), ("weight_bytes", ctypes.c_size_t),
                ("tile_cols", ctypes.c_int32), ("in_mode", ctypes.c_int32),
                ("B", ctypes.c_int32), ("cols_per_cloud", ctypes.c_int32),
                ("a_src", ctypes.c_void_p), ("a_ch", ctypes.c_int32), ("a_rows", ctypes.c_int32),
                ("idx", ctypes.c_void_p), ("xyz", ctypes.c_void_p), ("centers", ctypes.c_void_p),
                ("group_k", ctypes.c_int32),
                ("b_src", ctypes.c_void_p), ("b_ch", ctypes.c_int32), ("b_rows", ctypes.c_int32),
                ("nn_w", ctypes.c_void_p),
                ("out_mode", ctypes.c_int32), ("out", ctypes.c_void_p), ("ldo", ctypes.c_int32),
                ("pool_g", ctypes.c_int32), ("split_cout", ctypes.c_int32),
                ("l0_w", ctypes.c_void_p), ("l0_b", ctypes.c_void_p), ("l0_cout", ctypes.c_int32),
                ("in_bias", ctypes.c_void_p), ("out_prezeroed", ctypes.c_int32),
                ("win_cols", ctypes.c_int32), ("win_off", ctypes.c_int32), ("max_ctas", ctypes.c_int32),
                ("xyz_w", ctypes.c_void_p)]


def available():
    """The fused path needs the library and a CUDA device; there is no fallback inside it."""
    return torch.cuda.is_available() and hasattr(_lib.lib(), "cpfn_mlp_chain")


# ---- weight preparation (load time) -------------------------------------------------------------

def fold_bn(conv_w, conv_b, bn):
    """conv (1x1) followed by eval-mode BatchNorm == one affine map: W' = s W, b' = s (b - mean) + beta,
    s = gamma / sqrt(var + eps)."""
    w = conv_w.detach().double().reshape(conv_w.shape[0], conv_w.shape[1]).cpu()
    b = conv_b.detach().double().cpu()
    if bn is None:
        return w.float().numpy(), b.float().numpy()
    s = bn.weight.detach().double().cpu() / torch.sqrt(bn.running_var.detach().double().cpu() + bn.eps)
    return ((w * s[:, None]).float().numpy(),
            ((b - bn.running_mean.detach().double().cpu()) * s + bn.bias.detach().double().cpu()).float().numpy())


def pack_weights(w):
    """[cout, cin] float32 numpy -> uint8 numpy blob in the kernel's shared-memory image."""
    L = _lib.lib()
    w = np.ascontiguousarray(w, dtype=np.float32)
    cout, cin = w.shape
    nbytes = L.cpfn_mlp_packed_bytes(cout, cin)
    blob = np.zeros(nbytes, dtype=np.uint8)
    _lib.check(L.cpfn_mlp_pack_weights_host(w.ctypes.data, cout, cin, blob.ctypes.data), "mlp_pack_weights")
    return blob


def pad_bias(b):
    cout = b.shape[-1]
    pad = (cout + 127) // 128 * 128
    out = np.zeros(b.shape[:-1] + (pad,), dtype=np.float32)
    out[..., :cout] = b
    return out


class PackedChain:
    """Device-resident weights / biases of one chain: list of (W [cout,cin], b [cout], relu)."""

    def __init__(self, layers, device):
        self.dims = [(w.shape[1], w.shape[0], bool(r)) for w, _, r in layers]
        blob = np.concatenate([pack_weights(w) for w, _, _ in layers])
        self.weights = torch.from_numpy(blob).to(device)
        self.biases = [torch.from_numpy(pad_bias(b)).to(device) for _, b, _ in layers]
        self._layers = layers
        self._single = None

    def single_layer_chains(self):
        """The same layers as one-layer chains (for the layer-by-layer, channel-split execution of
        chains with few columns and wide layers: SA3, FP1)."""
        if self._single is None:
            self._single = [PackedChain([l], self.weights.device) for l in self._layers]
        return self._single


def run_chain(pc, B, cols_per_cloud, out, ldo, tile_cols=128, in_mode=IN_DENSE, a_src=None, a_ch=0, a_rows=0,
              idx=None, xyz=None, centers=None, group_k=0, b_src=None, b_ch=0, b_rows=0, nn_w=None,
              out_mode=OUT_ROWS, pool_g=0, biases=None, bias_per_cloud=(), masks=None, out_cm=None, split_cout=False,
              l0=None, in_bias=None, out_prezeroed=False, window=None, max_ctas=0, xyz_w=None):
    """Enqueue one fused chain on torch's current stream.  ``window`` = (first column, columns) of every cloud."""
    c = _Chain()
    c.n_layers = len(pc.dims)
    keep = []
    for l, (cin, cout, relu) in enumerate(pc.dims):
        bias = biases[l] if (biases is not None and biases[l] is not None) else pc.biases[l]
        keep.append(bias)
        c.layers[l].cin, c.layers[l].cout, c.layers[l].relu = cin, cout, int(relu)
        c.layers[l].bias_per_cloud = int(l in bias_per_cloud)
        c.layers[l].bias = bias.data_ptr()
        if masks and masks.get(l) is not None:          # (keep bits [cols, ceil(cout/32)] int32, scale of the kept values)
            c.layers[l].mask_bits, c.layers[l].mask_scale = masks[l][0].data_ptr(), float(masks[l][1])
        c.layers[l].out_cm = out_cm[l].data_ptr() if (out_cm and out_cm.get(l) is not None) else None
    c.weights, c.weight_bytes = pc.weights.data_ptr(), pc.weights.numel()
    c.tile_cols, c.in_mode, c.B, c.cols_per_cloud = tile_cols, in_mode, B, cols_per_cloud
    p = lambda t: t.data_ptr() if t is not None else None
    c.a_src, c.a_ch, c.a_rows = p(a_src), a_ch, a_rows
    c.idx, c.xyz, c.centers, c.group_k = p(idx), p(xyz), p(centers), group_k
    c.b_src, c.b_ch, c.b_rows, c.nn_w = p(b_src), b_ch, b_rows, p(nn_w)
    c.out_mode, c.out, c.ldo, c.pool_g = out_mode, out.data_ptr(), ldo, pool_g
    c.split_cout = int(split_cout)
    if l0 is not None:
        c.l0_w, c.l0_b, c.l0_cout = l0[0].data_ptr(), l0[1].data_ptr(), l0[0].shape[0]
    if in_bias is not None:
        c.in_bias = in_bias.data_ptr()
    c.out_prezeroed = int(bool(out_prezeroed))
    if window is not None:
        c.win_off, c.win_cols = int(window[0]), int(window[1])
    c.max_ctas = int(max_ctas)
    if xyz_w is not None:       # GROUP + features: the position columns of layer 0 as an epilogue term ([cout_pad, 4] fp32)
        c.xyz_w = xyz_w.data_ptr()
    with torch.cuda.device(out.device):
        _lib.check(_lib.lib().cpfn_mlp_chain(ctypes.byref(c), torch.cuda.current_stream(out.device).cuda_stream),
                   "mlp_chain")
    cuda_ops.count_launches(1)
    return out


def pick_tile(dims, cols_per_cloud, need_cloud_aligned=False, prefer=128, name=None):
    """Tile choice; CPFN_TILE_<SA1|SA2|HEAD>=128|64|32 overrides it (tuning experiments)."""
    if name and os.environ.get("CPFN_TILE_" + name):
        return int(os.environ["CPFN_TILE_" + name])
    return _pick_tile(dims, cols_per_cloud, need_cloud_aligned, prefer)


def _pick_tile(dims, cols_per_cloud, need_cloud_aligned=False, prefer=128):
    """Largest tile (128 / 64 / 32 columns) whose two activation buffers plus a 3-stage weight
    ring fit the 227 KB of shared memory (mirrors launch_chain in csrc/mlp_chain.cu)."""
    for tile in (128, 64, 32):
        if tile > prefer:
            continue
        if need_cloud_aligned and tile != 128 and cols_per_cloud % tile:
            continue                      # the channels-as-M kernel indexes masks / per-cloud biases per tile
        need = max((cin + 63) // 64 * 2 * tile * 128 for cin, _, _ in dims)   # one in-place activation buffer
        if tile == 128:
            # points-as-M kernel: two 128-column sub-tiles per CTA, <= 256 TMEM columns each
            if any((cout + 127) // 128 > 2 for _, cout, _ in dims):
                continue
            need *= 2
        elif any((cout + 127) // 128 > 512 // tile for _, cout, _ in dims[:-1]):
            continue
        if need + 1024 + 10240 + 3 * 16384 <= 227 * 1024:
            return tile
    raise RuntimeError("cpfn_b200.fused: chain does not fit shared memory: %r" % (dims,))


def run_layerwise(pc, B, cols_per_cloud, out, first, out_mode=OUT_ROWS, pool_g=0, bias0=None, out_prezeroed=False):
    """Few columns, wide layers: run the chain one layer per launch with one CTA per (64-column tile,
    128-channel chunk) so that the weight stream is spread over many SMs; activations between the
    layers are point-major fp32 rows in HBM (a few MB, L2 resident).  `first` = kwargs of the input
    (in_mode, a_src, ...) for the first layer."""
    dev = out.device
    chains = pc.single_layer_chains()
    x_kwargs = dict(first)
    for l, c in enumerate(chains):
        last = l == len(chains) - 1
        cout = c.dims[0][1]
        dst = out if last else torch.empty(B, cols_per_cloud, cout, dtype=torch.float32, device=dev)
        extra = {}
        if l == 0 and bias0 is not None:
            extra = dict(biases=[bias0], bias_per_cloud=(0,))
        tile = 64 if cols_per_cloud % 64 == 0 else 32
        if (c.dims[0][0] + 63) // 64 * 2 * tile * 128 + 5120 + 3 * 16384 > 227 * 1024:
            tile = 32
        if os.environ.get("CPFN_TILE_LAYERWISE"):
            tile = int(os.environ["CPFN_TILE_LAYERWISE"])
        run_chain(c, B, cols_per_cloud, dst, cout, tile_cols=tile, split_cout=True,
                  out_mode=(out_mode if last else OUT_ROWS), pool_g=(pool_g if last else 0),
                  out_prezeroed=(out_prezeroed and last), **x_kwargs, **extra)
        x_kwargs = dict(in_mode=IN_DENSE, a_src=dst, a_ch=cout, a_rows=cols_per_cloud)
    return out


def linear_rows(x, W, bias, out):
    """out[r, :cout] = bias + x[r] @ W.T in fp32 (x [rows, cin], W [cout, cin])."""
    rows, cin = x.shape
    cout = W.shape[0]
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().cpfn_linear_rows(x.data_ptr(), W.data_ptr(), bias.data_ptr(), rows, cin, cout,
                                               out.shape[1], out.data_ptr(), _stream(x)), "linear_rows")
    cuda_ops.count_launches(1)
    return out


def mlp_chain(*args, **kwargs):
    """Timed alias used by bench.py's per-op profile."""
    return run_chain(*args, **kwargs)


# ---- small fused helpers (csrc/glue.cu) ---------------------------------------------------------

def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def three_nn_weights(unknown, known, sorted_queries=None):
    """unknown [B,n,3], known [B,m,3] -> (weights f32 [B,n,3], idx int32 [B,n,3]): three_nn + sqrt +
    normalised inverse-distance weights in one kernel.  ``sorted_queries``: a ball-query grid workspace of ``unknown``
    (cpfn_ball_query_grid_build) -- the queries are then answered in its cell order (same result, coherent warps)."""
    B, n, _ = unknown.shape
    m = known.shape[1]
    w = torch.empty(B, n, 3, dtype=torch.float32, device=unknown.device)
    idx = torch.empty(B, n, 3, dtype=torch.int32, device=unknown.device)
    with torch.cuda.device(unknown.device):
        if sorted_queries is not None and 384 <= m <= 2048 and os.environ.get("CPFN_NN_SORTED", "1") != "0":
            _lib.check(_lib.lib().cpfn_three_nn_weights_sorted(sorted_queries.data_ptr(), known.data_ptr(), B, n, m,
                                                               w.data_ptr(), idx.data_ptr(), _stream(unknown)),
                       "three_nn_weights_sorted")
        else:
            _lib.check(_lib.lib().cpfn_three_nn_weights(unknown.data_ptr(), known.data_ptr(), B, n, m,
                                                        w.data_ptr(), idx.data_ptr(), _stream(unknown)), "three_nn_weights")
    cuda_ops.count_launches(1)
    return w, idx


def gather_xyz(xyz, idx):
    """xyz [B,N,3], idx int32 [B,S] -> [B,S,3]."""
    B, N, _ = xyz.shape
    S = idx.shape[1]
    out = torch.empty(B, S, 3, dtype=torch.float32, device=xyz.device)
    with torch.cuda.device(xyz.device):
        _lib.check(_lib.lib().cpfn_gather_xyz(xyz.data_ptr(), idx.data_ptr(), B, N, S, out.data_ptr(), _stream(xyz)),
                   "gather_xyz")
    cuda_ops.count_launches(1)
    return out


def spfn_post(heads, x_off, w_off, K, t_off=None, n_types=0, scatter=None):
    """heads [B,N,ld] contiguous -> (X [B,N,3] unit normals, W [B,N,K] softmax memberships,
    instance int32 [B,N] = argmax W, type int32 [B,N] = argmax of the type logits | None).
    ``scatter`` = dict(X=, W=, T=, stride=, offset=): the patch-sharded cascade's exchange fused into this kernel --
    X / W and the type logits T of local patch j are written into the slabs j * stride + offset of the given
    [n_patches, N, .] arrays, which may live in ANOTHER GPU's memory (peer-mapped, see api.LocalSPFN); the returned
    X / W are then those destination arrays."""
    B, N, ld = heads.shape
    dev = heads.device
    inst = torch.empty(B, N, dtype=torch.int32, device=dev)
    typ = torch.empty(B, N, dtype=torch.int32, device=dev) if t_off is not None else None
    if scatter is None:
        X = torch.empty(B, N, 3, dtype=torch.float32, device=dev)
        W = torch.empty(B, N, K, dtype=torch.float32, device=dev)
        T_out, geom = None, (0, 0, 0)
    else:
        X, W, T_out = scatter["X"], scatter["W"], scatter["T"]
        geom = (N, int(scatter["stride"]), int(scatter["offset"]))
        if N % 256 or X.shape[1:] != (N, 3) or W.shape[1:] != (N, K) or T_out.shape[1:] != (N, n_types) or \
                not (X.is_contiguous() and W.is_contiguous() and T_out.is_contiguous()) or \
                (B - 1) * geom[1] + geom[2] >= W.shape[0]:
            raise ValueError("spfn_post: scatter destination does not match the batch")
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().cpfn_spfn_post_scatter(heads.data_ptr(), B * N, ld, x_off, t_off or 0, n_types, w_off, K,
                                                     X.data_ptr(), W.data_ptr(),
                                                     T_out.data_ptr() if T_out is not None else None, inst.data_ptr(),
                                                     typ.data_ptr() if typ is not None else None, *geom,
                                                     _stream(heads)), "spfn_post")
    cuda_ops.count_launches(1)
    return X, W, inst, typ


# ---- PointNet2 on the fused chains --------------------------------------------------------------

_CACHE_ATTR = "_cpfn_fused_cache"


def invalidate(model):
    """Drop the folded / packed weights cached on the modules.  Not needed for correctness: every cache entry
    carries the signature of the tensors it was built from and is rebuilt when one of them changed."""
    for m in model.modules():
        for k in [k for k in vars(m) if k.startswith(_CACHE_ATTR)]:
            delattr(m, k)


def _conv_bn_sources(convs, bns):
    out = []
    for conv, bn in zip(convs, bns):
        out += [conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var]
    return out


def signature(tensors):
    """Changes whenever one of `tensors` is written in place (optimizer step, load_state_dict, BatchNorm's
    running statistics in train mode), moved or replaced: (storage address, version counter) per tensor."""
    return tuple((t.data_ptr(), t._version) for t in tensors)


def model_sources(model):
    """Every parameter / buffer the fused forward of a PointNet2 folds into its packed weights."""
    out = []
    for sa in (model.sa1, model.sa2, model.sa3):
        out += _conv_bn_sources(sa.conv_blocks[0], sa.bn_blocks[0])
    for fp in (model.sfp1, model.sfp2, model.sfp3):
        out += _conv_bn_sources(fp.mlp_convs, fp.mlp_bns)
    out += [model.fc1.weight, model.fc1.bias]
    if not model.features_extractor:
        out += [model.bn1.weight, model.bn1.bias, model.bn1.running_mean, model.bn1.running_var]
        for fc in model.fc2:
            out += [fc.weight, fc.bias]
    return out


def _cached(module, key, sources, device):
    """The cache entry `key` of `module` if it was built from the current values of `sources` on `device`."""
    entry = getattr(module, key, None)
    if entry is not None and entry[1] == device and entry[2] == signature(sources):
        return entry[0]
    return None


def _store(module, key, value, sources, device):
    setattr(module, key, (value, device, signature(sources)))
    return value


def _sa_chain(module, device):
    sources = _conv_bn_sources(module.conv_blocks[0], module.bn_blocks[0])
    pc = _cached(module, _CACHE_ATTR, sources, device)
    if pc is None:
        layers = []
        for j, (conv, bn) in enumerate(zip(module.conv_blocks[0], module.bn_blocks[0])):
            w, b = fold_bn(conv.weight, conv.bias, bn)
            if j == 0 and module.group_all and w.shape[1] > 3:
                # reference group_all order is [xyz, feats] (pointset_abstraction.py:54-56); the
                # kernel builds rows as [feats, xyz]: permute the input channels of the first layer
                w = np.concatenate([w[:, 3:], w[:, :3]], axis=1)
            layers.append((w, b, True))
        l0 = None
        if layers[0][0].shape[1] == 3 and layers[0][0].shape[0] % 4 == 0 and len(layers) > 1 and not module.group_all:
            # bare positions: the 3 -> C first layer runs in fp32 on the CUDA cores inside the tile builder
            w0, b0, _ = layers.pop(0)
            l0 = (torch.from_numpy(np.ascontiguousarray(w0)).to(device), torch.from_numpy(np.ascontiguousarray(b0)).to(device))
        pc = PackedChain(layers, device)
        pc.l0 = l0
        pc.alt = None
        w0, b0, _ = layers[0]
        D = w0.shape[1] - 3
        if l0 is None and not module.group_all and D > 0 and D % 64 == 0 and len(layers) > 1:
            # features + positions: the three position columns of layer 0 leave the tensor-core operand (they would cost a
            # whole 64-channel K atom of shared memory, which is what keeps 128-column tiles from fitting) and come back
            # as an fp32 term of layer 0's epilogue: rows (wx, wy, wz, bias) per output channel (cpfn_mlp_chain_t.xyz_w)
            rows = np.zeros(((w0.shape[0] + 127) // 128 * 128, 4), dtype=np.float32)
            rows[:w0.shape[0], :3] = w0[:, D:D + 3]
            rows[:w0.shape[0], 3] = b0
            alt = PackedChain([(np.ascontiguousarray(w0[:, :D]), b0, True)] + layers[1:], device)
            alt.xyz_w = torch.from_numpy(rows).to(device)
            pc.alt = alt
        _store(module, _CACHE_ATTR, pc, sources, device)
    return pc


def sa_indices(module, xyz):
    """FPS -> centroid gather -> ball query of a (non group_all) SA layer: (new_xyz [B,S,3], group_idx
    int32 [B,S,K]).  Depends on positions only, so it can run ahead of the previous layer's MLP."""
    _, new_xyz = cuda_ops.farthest_point_sampling(xyz, module.num_points, return_centroids=True)
    return new_xyz, cuda_ops.ball_query(new_xyz, xyz, module.radius_list[0], module.num_samples_list[0])


def sa_indices_overlapped(module, xyz, side, return_grid=False):
    """``sa_indices`` with the ball query's uniform grid -- which depends on the cloud and the radius, not on the
    centroids -- built on ``side`` while farthest point sampling runs on the current stream (FPS leaves more than
    half of the SMs idle).  Same result as ``sa_indices``."""
    B, N, _ = xyz.shape
    dev = xyz.device
    radius, K = float(module.radius_list[0]), int(module.num_samples_list[0])
    L = _lib.lib()
    if side is None or not (2048 <= N <= 32768) or radius <= 0 or os.environ.get("CPFN_BQ_NO_GRID"):
        return sa_indices(module, xyz) + ((None,) if return_grid else ())
    main = torch.cuda.current_stream(dev)
    nbytes = L.cpfn_ball_query_grid_workspace_bytes(B, N)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    fork = torch.cuda.Event()
    fork.record(main)
    with torch.cuda.stream(side), torch.cuda.device(dev):
        side.wait_event(fork)
        _lib.check(L.cpfn_ball_query_grid_build(xyz.data_ptr(), B, N, radius, ws.data_ptr(), nbytes, side.cuda_stream),
                   "ball_query_grid_build")
        built = torch.cuda.Event()
        built.record(side)
    ws.record_stream(side)
    _, new_xyz = cuda_ops.farthest_point_sampling(xyz, module.num_points, return_centroids=True)   # centroids from the kernel
    S = new_xyz.shape[1]
    idx = torch.empty(B, S, K, dtype=torch.int32, device=dev)
    main.wait_event(built)
    with torch.cuda.device(dev):
        _lib.check(L.cpfn_ball_query_grid_query(new_xyz.data_ptr(), xyz.data_ptr(), B, N, S, radius, K, idx.data_ptr(),
                                                ws.data_ptr(), nbytes, main.cuda_stream), "ball_query_grid_query")
    cuda_ops.count_launches(2)
    if return_grid:
        return new_xyz, idx, ws
    return new_xyz, idx


def sa1_pipelined(module, xyz, worker, n_chunks=4):
    """Set abstraction on bare positions with the sampling, the ball query and the MLP PIPELINED over chunks of
    centroids.  Furthest point sampling is a chain of dependent rounds that keeps fewer than half of the SMs busy
    (16 clouds x 4 CTAs); its centroids come out in order, so the sampling runs as ``n_chunks`` launches
    (cpfn_furthest_point_sampling_rounds, bit-identical to one) on the current stream and, as soon as a chunk of
    centroids exists, their ball query (cpfn_ball_query_grid_query_range) and their column window of the fused MLP
    chain run on the stream ``worker`` -- on the SMs the sampling leaves idle -- while the next chunk is being
    sampled.  Returns (new_xyz [B,S,3], new_feats_pm [B,S,D'], event): wait for the event before reading new_feats_pm;
    new_xyz is complete on the current stream.  None when the shapes do not allow it (caller falls back)."""
    B, N, _ = xyz.shape
    dev = xyz.device
    L = _lib.lib()
    S, K = int(module.num_points), int(module.num_samples_list[0])
    radius = float(module.radius_list[0])
    pc = _sa_chain(module, dev)
    tile = pick_tile(pc.dims, S * K, name='SA1', prefer=128)
    if (n_chunks < 2 or S % n_chunks or ((S // n_chunks) * K) % tile or K % 32 or not (2048 <= N <= 16384) or radius <= 0
            or getattr(pc, "l0", None) is None or not L.cpfn_fps_rounds_supported(B, N) or os.environ.get("CPFN_BQ_NO_GRID")):
        return None
    main = torch.cuda.current_stream(dev)
    cout = pc.dims[-1][1]
    fps_idx = torch.empty(B, S, dtype=torch.int32, device=dev)
    new_xyz = torch.empty(B, S, 3, dtype=torch.float32, device=dev)
    state = torch.empty(B, N, dtype=torch.float32, device=dev)
    gidx = torch.empty(B, S, K, dtype=torch.int32, device=dev)
    out = torch.empty(B, S, cout, dtype=torch.float32, device=dev)
    nbytes = L.cpfn_ball_query_grid_workspace_bytes(B, N)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    floor = int(os.environ.get("CPFN_FPS_FLOOR_KB", "120")) * 1024
    # while the sampling runs, the chain may only use the SMs it leaves free (two CTAs each): no CTA of a chunk's
    # chain is left pending that could slip onto the sampling SMs between two sampling launches
    sms = L.cpfn_sm_count()
    cap = int(os.environ.get("CPFN_SA1_CAP", "0")) or max(2, 2 * (sms - 4 * B))
    fork = torch.cuda.Event()
    fork.record(main)
    with torch.cuda.stream(worker), torch.cuda.device(dev):
        worker.wait_event(fork)
        _lib.check(L.cpfn_zero_fill(out.data_ptr(), out.numel() * 4, worker.cuda_stream), "zero_fill")
        _lib.check(L.cpfn_ball_query_grid_build(xyz.data_ptr(), B, N, radius, ws.data_ptr(), nbytes, worker.cuda_stream),
                   "ball_query_grid_build")
    per = S // n_chunks
    for c in range(n_chunks):
        j0, j1 = c * per, (c + 1) * per
        with torch.cuda.device(dev):
            _lib.check(L.cpfn_furthest_point_sampling_rounds(xyz.data_ptr(), B, N, S, j0, j1, fps_idx.data_ptr(),
                                                             new_xyz.data_ptr(), state.data_ptr(), state.numel() * 4,
                                                             floor, main.cuda_stream), "furthest_point_sampling_rounds")
        sampled = torch.cuda.Event()
        sampled.record(main)
        with torch.cuda.stream(worker), torch.cuda.device(dev):
            worker.wait_event(sampled)
            _lib.check(L.cpfn_ball_query_grid_query_range(new_xyz.data_ptr(), xyz.data_ptr(), B, N, S, j0, per, radius, K,
                                                          gidx.data_ptr(), ws.data_ptr(), nbytes, worker.cuda_stream),
                       "ball_query_grid_query_range")
            run_chain(pc, B, S * K, out, cout, tile_cols=tile, in_mode=IN_GROUP, a_src=None, a_ch=0, a_rows=N, idx=gidx,
                      xyz=xyz, centers=new_xyz, group_k=K, out_mode=OUT_POOL, pool_g=K, l0=pc.l0, out_prezeroed=True,
                      window=(j0 * K, per * K), max_ctas=(cap if c + 1 < n_chunks else 0))
    done = torch.cuda.Event()
    done.record(worker)
    for t in (fps_idx, new_xyz, state, gidx, out, ws):
        t.record_stream(worker)
    cuda_ops.count_launches(2 + 2 * n_chunks)
    return new_xyz, out, done


def sa_forward_pm(module, xyz, feats_pm, indices=None, out=None):
    """Point-major set abstraction.  xyz [B,N,3], feats_pm [B,N,D] | None ->
    (new_xyz [B,S,3] | None, new_feats_pm [B,S,D']).  ``out``: a ZERO-FILLED [B,S,D'] buffer for the pooled
    features (the caller filled it off the critical path; the chain then skips its own memset)."""
    assert len(module.radius_list) == 1, "multi-scale grouping: use the per-op path"
    B, N, _ = xyz.shape
    dev = xyz.device
    pc = _sa_chain(module, dev)
    cout = pc.dims[-1][1]
    D = 0 if feats_pm is None else feats_pm.shape[2]
    if module.group_all:
        idx, centers = _group_all_constants(B, N, dev)        # every point of the cloud, centre at the origin
        pre = out is not None
        if not pre:
            out = torch.empty(B, 1, cout, dtype=torch.float32, device=dev)
        first = dict(in_mode=IN_GROUP, a_src=feats_pm, a_ch=D, a_rows=N, idx=idx, xyz=xyz, centers=centers, group_k=N)
        if N % 32 == 0 and B * N <= 16384:
            run_layerwise(pc, B, N, out, first, out_mode=OUT_POOL, pool_g=N, out_prezeroed=pre)
        else:
            tile = pick_tile(pc.dims, N, need_cloud_aligned=True, prefer=32 if cout > 512 else 128)
            run_chain(pc, B, N, out, cout, tile_cols=tile, out_mode=OUT_POOL, pool_g=N, out_prezeroed=pre, **first)
        return None, out
    S, K = module.num_points, module.num_samples_list[0]
    if indices is None:
        indices = sa_indices(module, xyz)
    new_xyz, group_idx = indices
    pre = out is not None
    if not pre:
        out = torch.empty(B, S, cout, dtype=torch.float32, device=dev)
    alt = getattr(pc, "alt", None)
    if (alt is not None and D > 0 and (S * K) % 128 == 0 and K in (32, 64, 128) and cout <= 256
            and os.environ.get("CPFN_SA_XYZ_EPILOGUE", "1") != "0"
            and max((cin + 63) // 64 for cin, _, _ in alt.dims) * 2 * 128 * 128 + 12288 + 2 * 16384 <= 113 * 1024):
        # 128-column tiles on the points-as-M kernel (half the epilogue instructions per element of the 64-column,
        # channels-as-M kernel), positions as an epilogue term, pooled through shared memory
        run_chain(alt, B, S * K, out, cout, tile_cols=128, in_mode=IN_GROUP, a_src=feats_pm, a_ch=D, a_rows=N,
                  idx=group_idx, xyz=xyz, centers=new_xyz, group_k=K, out_mode=OUT_POOL, pool_g=K, xyz_w=alt.xyz_w,
                  out_prezeroed=pre)
        return new_xyz, out
    run_chain(pc, B, S * K, out, cout, tile_cols=pick_tile(pc.dims, S * K, name=('SA1' if D == 0 else 'SA2'), prefer=(128 if D == 0 else 64)), in_mode=IN_GROUP, a_src=feats_pm, a_ch=D, a_rows=N,
              idx=group_idx, xyz=xyz, centers=new_xyz, group_k=K, out_mode=OUT_POOL, pool_g=K, l0=getattr(pc, "l0", None),
              out_prezeroed=pre)
    return new_xyz, out


def _fp_chain(module, device, split=None, tail=None):
    """split = number of leading input channels of layer 0 that are a per-cloud constant
    (FP1: the broadcast global feature) -> returns (chain without them, const-part weight).
    tail = (W [cout,cin] numpy, tag, the tensors W was folded from): one more layer without bias / ReLU appended to the chain -- the linear part
    of the NEXT module's first layer, applied here once per source row (see pointnet2_forward)."""
    key = _CACHE_ATTR + ("_s%d_%d" % split if split else "") + ("_t" + tail[1] if tail else "")
    sources = _conv_bn_sources(module.mlp_convs, module.mlp_bns) + (list(tail[2]) if tail else [])
    pc = _cached(module, key, sources, device)
    if pc is None:
        layers, wconst = [], None
        for j, (conv, bn) in enumerate(zip(module.mlp_convs, module.mlp_bns)):
            w, b = fold_bn(conv.weight, conv.bias, bn)
            if j == 0 and split:
                # the constant part becomes its own one-layer chain (with the layer's bias)
                wconst = (torch.from_numpy(np.ascontiguousarray(w[:, split[0]:split[1]])).to(device),
                          torch.from_numpy(np.ascontiguousarray(b)).to(device))
                w = np.concatenate([w[:, :split[0]], w[:, split[1]:]], axis=1)
            layers.append((w, b, True))
        if tail is not None:
            layers.append((np.ascontiguousarray(tail[0], dtype=np.float32), np.zeros(tail[0].shape[0], np.float32), False))
        pc = _store(module, key, (PackedChain(layers, device), wconst), sources, device)
    return pc


def fp_forward_pm(module, xyz1, xyz2, feats1_pm, feats2_pm, nn=None, tail=None):
    """Point-major feature propagation.  xyz1 [B,N,3], xyz2 [B,S,3] | None, feats1_pm [B,N,D1] | None,
    feats2_pm [B,S,D2] -> [B,N,D']."""
    B, N, _ = xyz1.shape
    dev = xyz1.device
    D1 = 0 if feats1_pm is None else feats1_pm.shape[2]
    if xyz2 is None:
        # pos2 is None: feats2 is one row per cloud, repeated over the N points (reference :33-34).
        # Its contribution to layer 0 is a per-cloud constant: fold it into a per-cloud bias.
        D2 = feats2_pm.shape[2]
        pc, wconst = _fp_chain(module, dev, split=(D1, D1 + D2))
        g = feats2_pm.reshape(B, D2).contiguous()
        # per-cloud bias rows, allocated per call (a captured graph keeps its own copy alive): linear_rows writes the
        # cout valid entries and the zero padding
        bias0 = torch.empty(B, pc.biases[0].numel(), dtype=torch.float32, device=dev)
        linear_rows(g, wconst[0], wconst[1], bias0)          # fp32: W_global @ g + b, one vector per cloud
        out = torch.empty(B, N, pc.dims[-1][1], dtype=torch.float32, device=dev)
        first = dict(in_mode=IN_DENSE, a_src=feats1_pm, a_ch=D1, a_rows=N)
        if N % 32 == 0 and B * N <= 16384:
            run_layerwise(pc, B, N, out, first, bias0=bias0)
        else:
            tile = pick_tile(pc.dims, N, need_cloud_aligned=True)
            run_chain(pc, B, N, out, out.shape[2], tile_cols=tile, biases=[bias0] + [None] * (len(pc.dims) - 1),
                      bias_per_cloud=(0,), **first)
        return out
    pc, _ = _fp_chain(module, dev, tail=tail)
    w, idx = nn if nn is not None else three_nn_weights(xyz1, xyz2)
    out = torch.empty(B, N, pc.dims[-1][1], dtype=torch.float32, device=dev)
    run_chain(pc, B, N, out, out.shape[2], tile_cols=pick_tile(pc.dims, N), in_mode=IN_INTERP, a_src=feats1_pm, a_ch=D1, a_rows=N,
              idx=idx, b_src=feats2_pm, b_ch=feats2_pm.shape[2], b_rows=feats2_pm.shape[1], nn_w=w)
    return out


_group_all_cache = {}


def _group_all_constants(B, N, dev):
    """(group index 0..N-1 per cloud int32 [B*N], zero centres [B,3]) of a group_all SA layer -- constants,
    built once per shape instead of three tiny kernels on the critical path of every step."""
    key = (B, N, dev)
    c = _group_all_cache.get(key)
    if c is None:
        c = _group_all_cache[key] = (torch.arange(N, dtype=torch.int32, device=dev).repeat(B),
                                     torch.zeros(B, 3, dtype=torch.float32, device=dev))
    return c
_side_streams = {}
USE_SIDE_STREAM = True      # bench.py's per-op profiling pass serialises everything on one stream


def _side_stream(dev, which=0):
    key = (dev.type, dev.index, which)
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=dev)
    return _side_streams[key]


# ---- always-on dropout (pn2_network.py:63) as a bit mask drawn from torch's CUDA generator -----------------------

_rng_state = {}


def _rng_state_tensor(dev, slot=0):
    key = (dev.type, dev.index, slot)
    t = _rng_state.get(key)
    if t is None:
        t = _rng_state[key] = torch.zeros(2, dtype=torch.int64, device=dev)      # {seed, offset}, read by the mask kernel
    return t


def dropout_plan(nelem, dev):
    """(threads, generator advance) of the launch torch's F.dropout makes for `nelem` float32 elements on `dev`
    (ATen native/cuda/Dropout.cu: 256-thread blocks, at most SMs x max-threads/256 of them, 4 elements per
    thread and iteration; the generator moves on by 4 x iterations)."""
    props = torch.cuda.get_device_properties(dev)
    blocks = min(props.multi_processor_count * (props.max_threads_per_multi_processor // 256), (nelem + 255) // 256)
    threads = max(1, blocks) * 256
    return threads, ((max(nelem, 1) - 1) // (threads * 4) + 1) * 4


def sync_rng(dev, nelem, slot=0):
    """Hand the current (seed, offset) of torch's CUDA generator to the device-side RNG state the mask kernel
    reads, and advance the generator exactly as F.dropout on `nelem` elements would: torch.manual_seed governs the
    masks, and they are the ones the reference's own F.dropout call draws on this GPU.  Stream-ordered (a one-thread
    kernel); call it before every replay of a captured forward -- it cannot be part of the capture.  ``slot``
    selects one of several independent device-side states, for forwards that are in flight at the same time."""
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    gen = torch.cuda.default_generators[idx]
    seed, offset = gen.initial_seed(), gen.get_offset()
    gen.set_offset(offset + dropout_plan(nelem, dev)[1])
    state = _rng_state_tensor(dev, slot)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().cpfn_rng_set(state.data_ptr(), seed & 0xFFFFFFFFFFFFFFFF, offset,
                                           torch.cuda.current_stream(dev).cuda_stream), "rng_set")
    cuda_ops.count_launches(1)


def dropout_bits(B, C, N, dev, p=0.5, slot=0):
    """Keep bits of F.dropout(x [B,C,N], p): (int32 [B*N, ceil(C/32)], scale of the kept values).  Outside a graph
    capture the RNG state is taken from torch's generator here; a captured launch reads whatever ``sync_rng`` wrote
    before the replay."""
    if not torch.cuda.is_current_stream_capturing():
        sync_rng(dev, B * C * N, slot)
    words = (C + 31) // 32
    bits = torch.empty(B * N, words, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().cpfn_dropout_mask_bits(_rng_state_tensor(dev, slot).data_ptr(), B, C, N, 1.0 - p,
                                                     dropout_plan(B * C * N, dev)[0], bits.data_ptr(),
                                                     torch.cuda.current_stream(dev).cuda_stream), "dropout_mask_bits")
    cuda_ops.count_launches(1)
    return bits, 1.0 / (1.0 - p)


def _pm(t):
    return None if t is None else t.permute(0, 2, 1).contiguous()


def set_abstraction_forward(module, pos, feats):
    """Reference-layout wrapper: pos [B,3,N], feats [B,D,N] | None -> (new_pos [B,3,S] | None, [B,D',S])."""
    new_xyz, out = sa_forward_pm(module, _pm(pos.float()), _pm(feats))
    return (None if new_xyz is None else new_xyz.permute(0, 2, 1)), out.permute(0, 2, 1)


def feature_propagation_forward(module, pos1, pos2, feats1, feats2):
    out = fp_forward_pm(module, _pm(pos1.float()), _pm(pos2), _pm(feats1), _pm(feats2))
    return out.permute(0, 2, 1)


def _head_chain(model, device):
    """FP3 (from its second layer on) + fc1 + bn1 + ReLU + the heads as ONE chain.  FP3's input is the 3-NN
    interpolation of l5 alone (pn2_network.py:58: feats1 is None), so its first layer commutes with the
    interpolation: W1 (sum_j w_j f_j) = sum_j w_j (W1 f_j).  W1 (BatchNorm folded) is applied to the 512 rows of
    l5 per cloud as one more layer of FP2's chain instead of to the 8192 interpolated rows; the bias and the ReLU
    are applied where the rows are interpolated (``in_bias``).  Returns (chain, head sizes, W1, b1)."""
    sources = (_conv_bn_sources(model.sfp3.mlp_convs, model.sfp3.mlp_bns)
               + [model.fc1.weight, model.fc1.bias, model.bn1.weight, model.bn1.bias, model.bn1.running_mean,
                  model.bn1.running_var] + [t for fc in model.fc2 for t in (fc.weight, fc.bias)])
    pc = _cached(model, _CACHE_ATTR, sources, device)
    if pc is None:
        layers, first = [], None
        for j, (conv, bn) in enumerate(zip(model.sfp3.mlp_convs, model.sfp3.mlp_bns)):
            w, b = fold_bn(conv.weight, conv.bias, bn)
            if j == 0:
                first = (w, torch.from_numpy(np.ascontiguousarray(b, dtype=np.float32)).to(device))
            else:
                layers.append((w, b, True))
        w, b = fold_bn(model.fc1.weight, model.fc1.bias, model.bn1)
        layers.append((w, b, True))
        hw = np.concatenate([fold_bn(fc.weight, fc.bias, None)[0] for fc in model.fc2], axis=0)
        hb = np.concatenate([fold_bn(fc.weight, fc.bias, None)[1] for fc in model.fc2], axis=0)
        layers.append((hw, hb, False))
        pc = _store(model, _CACHE_ATTR, (PackedChain(layers, device), [fc.out_channels for fc in model.fc2], first[0],
                                         first[1], sources[:6]), sources, device)
    return pc


@torch.no_grad()
def pointnet2_forward(model, P, dropout=True, rng_slot=0):
    """Whole PointNet2 forward (reference pn2_network.py:38-73) on the fused kernels.
    P [B,N,3] float32 CUDA.  FP3, fc1 + bn1 + ReLU, dropout and the heads are ONE chain.
    Returns (heads [list of [B,N,o_i]], l3_feats [B,1024,1], output_feat [B,128,N], l1_xyz, l2_xyz)."""
    assert not (model.use_glob_features or model.use_loc_features or model.features_extractor)
    P = P.float().contiguous()
    B, N, _ = P.shape
    dev = P.device
    # Everything that depends on positions only (SA2's sampling / ball query, the 3-NN weights of FP2 and
    # FP3) runs on a side stream, concurrently with SA1's ball query and MLP chain on the main stream.
    main = torch.cuda.current_stream(dev)
    side = _side_stream(dev) if USE_SIDE_STREAM else main
    mask = None
    # SA1: sampling pipelined with its own ball query / MLP (sa1_pipelined) when the shapes allow it
    n_chunks = int(os.environ.get("CPFN_SA1_CHUNKS", "1"))      # off by default: measured slower
    piped = sa1_pipelined(model.sa1, P, _side_stream(dev, 1), n_chunks) if (side is not main and n_chunks > 1) else None
    if piped is not None:
        l1_xyz, l1, l1_done = piped
    else:
        # (the grid workspace starts with the cloud in cell order: FP3's 3-NN answers its queries in that order)
        *idx1, grid_ws = sa_indices_overlapped(model.sa1, P, side if side is not main else None, return_grid=True)
        l1_xyz = idx1[0]
    fork = torch.cuda.Event()
    fork.record(main)
    with torch.cuda.stream(side):
        side.wait_event(fork)
        idx2 = sa_indices(model.sa2, l1_xyz)
        nn3 = three_nn_weights(P, l1_xyz, sorted_queries=grid_ws if piped is None else None)
        nn2 = three_nn_weights(l1_xyz, idx2[0])
        # the reference's always-on dropout (pn2_network.py:63): same generator, same mask, as 1 bit per element.
        # (Not under the sampling: its CTAs would share the sampling SMs and stretch every round -- measured.)
        if dropout:
            mask = dropout_bits(B, 128, N, dev, p=0.5, slot=rng_slot)
        join = torch.cuda.Event()
        join.record(side)
    if piped is not None:
        main.wait_event(l1_done)
    else:
        _, l1 = sa_forward_pm(model.sa1, P, None, indices=idx1)
    main.wait_event(join)
    if side is not main:
        for t in (*idx2, *nn3, *nn2) + ((mask[0],) if mask is not None else ()):
            t.record_stream(main)
        if piped is None and grid_ws is not None:
            grid_ws.record_stream(side)
    l2_xyz, l2 = sa_forward_pm(model.sa2, l1_xyz, l1, indices=idx2)
    _, l3 = sa_forward_pm(model.sa3, l2_xyz, l2)
    l4 = fp_forward_pm(model.sfp1, l2_xyz, None, l2, l3)
    pc, head_sizes, w_fp3, b_fp3, fp3_src = _head_chain(model, dev)
    l5 = fp_forward_pm(model.sfp2, l1_xyz, l2_xyz, l1, l4, nn=nn2, tail=(w_fp3, "fp3", fp3_src))   # = W1_fp3 @ (FP2 output)
    n_out = sum(head_sizes)
    w, idx = nn3
    heads = torch.empty(B, N, n_out, dtype=torch.float32, device=dev)
    output_feat = torch.empty(B, 128, N, dtype=torch.float32, device=dev)
    fc1_layer = len(pc.dims) - 2
    masks = {fc1_layer: mask} if mask is not None else None
    run_chain(pc, B, N, heads, n_out, tile_cols=pick_tile(pc.dims, N, need_cloud_aligned=True, name='HEAD'), in_mode=IN_INTERP, a_src=None, a_ch=0, a_rows=N, idx=idx,
              b_src=l5, b_ch=l5.shape[2], b_rows=l5.shape[1], nn_w=w, masks=masks, out_cm={fc1_layer: output_feat},
              in_bias=b_fp3)
    outs, o = [], 0
    for n in head_sizes:
        outs.append(heads[:, :, o:o + n])
        o += n
    return outs, l3.reshape(B, -1, 1), output_feat, l1_xyz, l2_xyz, heads
