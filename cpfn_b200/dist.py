"""Patch / cloud sharding across the GPUs of one box.

Units (clouds of a GlobalSPFN batch, or the <= 32 patches of a LocalSPFN shape) are independent:
rank r takes units r, r+G, r+2G, ... with replicated weights and NO data-path collective.  The
single exchange step of the cascade is the all-gather of the per-patch records to the rank that
runs the merge (reference: evaluation_localSPFN.py:95-110 feeds Utils/merging_utils.py:6-52 with
every patch's W): one flat buffer, one collective (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_units(n_units, rank, world):
    """Indices of the units rank `rank` owns (round robin, as patch i -> rank i mod G)."""
    return list(range(rank, n_units, world))


def units_per_rank(n_units, world):
    return (n_units + world - 1) // world


def all_gather_patches(feats, indices, n_units, group=None):
    """The exchange step of the patch-sharded cascade on tensors: rank r holds its patches' per-point outputs
    ``feats`` f32 [b, Np, F] (memberships | normals | type logits, concatenated along F) and ``indices`` int [b, Np]
    (global point ids) for units r, r+G, ... in that order (b may be 0).  Returns (feats [n_units, Np, F], indices
    int64 [n_units, Np]) in UNIT order on every rank: one collective per dtype (``all_gather_into_tensor`` on NCCL --
    NVLink / NVSwitch on one box -- plain ``all_gather`` on gloo), then one row permutation."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    nccl = dist.get_backend(group) == "nccl"
    per = units_per_rank(n_units, world)
    mine = len(shard_units(n_units, rank, world))
    if feats.shape[0] != mine or indices.shape[0] != mine:
        raise ValueError("rank %d passes %d patches, its share of %d units is %d" % (rank, feats.shape[0], n_units, mine))
    if feats.dtype != torch.float32:
        raise TypeError("all_gather_patches: float32 features required")
    Np, F = feats.shape[1], feats.shape[2]            # every rank knows the record shape (it depends on the model only)
    dev = feats.device
    send = torch.zeros(per, Np, F, dtype=torch.float32, device=dev) if mine < per else feats.contiguous()
    send_idx = torch.zeros(per, Np, dtype=torch.int64, device=dev)
    if mine < per:
        send[:mine] = feats
    send_idx[:mine] = indices.to(torch.int64)
    recv = torch.empty(world * per, Np, F, dtype=torch.float32, device=dev)
    recv_idx = torch.empty(world * per, Np, dtype=torch.int64, device=dev)
    if nccl:
        dist.all_gather_into_tensor(recv, send, group=group)
        dist.all_gather_into_tensor(recv_idx, send_idx, group=group)
    else:
        dist.all_gather(list(recv.view(world, per, Np, F).unbind(0)), send, group=group)
        dist.all_gather(list(recv_idx.view(world, per, Np).unbind(0)), send_idx, group=group)
    order = torch.tensor([(u % world) * per + u // world for u in range(n_units)], dtype=torch.int64, device=dev)
    return recv.index_select(0, order), recv_idx.index_select(0, order)


def gather_patch_records(records, n_units, group=None):
    """All-gather the per-patch records of every rank; returns the list of `n_units` records in unit
    order (unit i lives on rank i mod G) on every rank.  A record is a dict with ``W`` f32 [n,K] memberships,
    ``X`` f32 [n,3] normals, ``T`` f32 [n,n_types] type logits, ``patch_indices`` int64 [n] and ``parameters``
    f32 [K,P] fitted parameters (P = 22 for the four primitive types; may be [K,0]).  Two collectives carry the
    payload: one for the float fields (one flat buffer), one for the int64 indices.  Ranks without patches (more
    ranks than patches) learn the record shape from an all-reduce."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    nccl = dist.get_backend(group) == "nccl"
    if records:
        dev = records[0]["W"].device
    elif nccl:
        dev = torch.device("cuda", torch.cuda.current_device())
    else:
        dev = torch.device("cpu")
    per = units_per_rank(n_units, world)
    if len(records) != len(shard_units(n_units, rank, world)):
        raise ValueError("rank %d holds %d records, its share of %d units is %d" %
                         (rank, len(records), n_units, len(shard_units(n_units, rank, world))))
    shapes = torch.zeros(4, dtype=torch.int64, device=dev)
    if records:
        r0 = records[0]
        n, K = r0["W"].shape
        shapes[0], shapes[1], shapes[2], shapes[3] = n, K, r0["T"].shape[-1], r0["parameters"].shape[-1]
        for r in records:
            if r["W"].dtype != torch.float32 or r["X"].dtype != torch.float32 or r["T"].dtype != torch.float32:
                raise TypeError("gather_patch_records: W, X and T must be float32")
            if (tuple(r["W"].shape) != (n, K) or tuple(r["X"].shape) != (n, 3) or r["T"].shape != r0["T"].shape
                    or r["patch_indices"].numel() != n or r["parameters"].shape != r0["parameters"].shape):
                raise ValueError("gather_patch_records: the records of one call must share their shapes")
    mine = shapes.clone()
    dist.all_reduce(shapes, op=dist.ReduceOp.MAX, group=group)      # ranks without patches learn the shapes
    if records and not torch.equal(mine, shapes):
        raise ValueError("gather_patch_records: ranks disagree on the record shape: %s vs %s" %
                         (mine.tolist(), shapes.tolist()))
    n, K, n_types, n_par = (int(v) for v in shapes.tolist())
    fields = (("W", n * K, (n, K)), ("X", n * 3, (n, 3)), ("T", n * n_types, (n, n_types)),
              ("parameters", K * n_par, (K, n_par)))
    rec_len = sum(size for _, size, _ in fields)
    flat = torch.zeros(per, rec_len, dtype=torch.float32, device=dev)
    flat_idx = torch.zeros(per, n, dtype=torch.int64, device=dev)
    for i, r in enumerate(records):
        flat[i] = torch.cat([r[name].reshape(-1).to(torch.float32) for name, _, _ in fields])
        flat_idx[i] = r["patch_indices"].reshape(-1).to(torch.int64)
    out = torch.empty(world, per, rec_len, dtype=torch.float32, device=dev)
    out_idx = torch.empty(world, per, n, dtype=torch.int64, device=dev)
    if nccl:
        dist.all_gather_into_tensor(out.view(world * per, rec_len), flat, group=group)
        dist.all_gather_into_tensor(out_idx.view(world * per, n), flat_idx, group=group)
    else:
        dist.all_gather(list(out.unbind(0)), flat, group=group)
        dist.all_gather(list(out_idx.unbind(0)), flat_idx, group=group)
    result = []
    for u in range(n_units):
        row = out[u % world, u // world]
        o = 0
        rec = {}
        for name, size, shape in fields:
            rec[name] = row[o:o + size].view(shape)
            o += size
        rec["patch_indices"] = out_idx[u % world, u // world]
        result.append(rec)
    return result
