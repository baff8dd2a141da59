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


def gather_patch_records(records, n_units, group=None):
    """All-gather the per-patch records of every rank; returns the list of `n_units` records in unit
    order (unit i lives on rank i mod G) on every rank.  One collective for the payload."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if records:
        dev = records[0]["W"].device
    elif dist.get_backend(group) == "nccl":
        dev = torch.device("cuda", torch.cuda.current_device())
    else:
        dev = torch.device("cpu")
    per = units_per_rank(n_units, world)
    shapes = torch.zeros(2, dtype=torch.int64, device=dev)
    if records:
        shapes[0], shapes[1] = records[0]["W"].shape
    dist.all_reduce(shapes, op=dist.ReduceOp.MAX, group=group)      # n, K (ranks without patches learn them)
    n, K = int(shapes[0]), int(shapes[1])
    rec_len = n * K + n * 3 + n * 4 + n * 2 + K * 22
    flat = torch.zeros(per, rec_len, dtype=torch.float32, device=dev)
    for i, r in enumerate(records):
        flat[i] = torch.cat([r["W"].reshape(-1), r["X"].reshape(-1), r["T"].reshape(-1),
                             r["patch_indices"].to(torch.int64).contiguous().view(torch.float32).reshape(-1),
                             r["parameters"].reshape(-1)])
    out = torch.empty(world, per, rec_len, dtype=torch.float32, device=dev)
    if dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(out.view(world * per, rec_len), flat, group=group)
    else:
        dist.all_gather(list(out.unbind(0)), flat, group=group)
    result = []
    for u in range(n_units):
        row = out[u % world, u // world]
        o = 0
        rec = {}
        for name, size, shape in (("W", n * K, (n, K)), ("X", n * 3, (n, 3)), ("T", n * 4, (n, 4))):
            rec[name] = row[o:o + size].view(shape)
            o += size
        rec["patch_indices"] = row[o:o + 2 * n].contiguous().view(torch.int64)
        o += 2 * n
        rec["parameters"] = row[o:o + K * 22].view(K, 22)
        result.append(rec)
    return result
