"""World-size-2 gloo tests (CPU) of the patch sharding and the one exchange step."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cpfn_b200 import dist as cd


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make_record(u, n=64, K=5):
    g = torch.Generator().manual_seed(100 + u)
    return {"W": torch.rand(n, K, generator=g), "X": torch.randn(n, 3, generator=g), "T": torch.randn(n, 4, generator=g),
            "patch_indices": torch.randint(0, 1 << 40, (n,), generator=g, dtype=torch.int64),
            "parameters": torch.randn(K, 22, generator=g)}


def _worker(rank, world, port, n_units, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = cd.shard_units(n_units, rank, world)
        recs = [_make_record(u) for u in mine]
        got = cd.gather_patch_records(recs, n_units)
        ok = len(got) == n_units
        for u in range(n_units):
            ref = _make_record(u)
            for k in ref:
                ok = ok and torch.equal(got[u][k], ref[k])
        q.put((rank, ok, mine))
    finally:
        dist.destroy_process_group()


def _patch_tensors(u, Np=48, F=28):
    g = torch.Generator().manual_seed(500 + u)
    return torch.randn(Np, F, generator=g), torch.randint(0, 1 << 40, (Np,), generator=g, dtype=torch.int64)


def _worker_tensors(rank, world, port, n_units, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = cd.shard_units(n_units, rank, world)
        if mine:
            feats = torch.stack([_patch_tensors(u)[0] for u in mine])
            idx = torch.stack([_patch_tensors(u)[1] for u in mine])
        else:
            feats, idx = torch.empty(0, 48, 28), torch.empty(0, 48, dtype=torch.int64)
        f_all, i_all = cd.all_gather_patches(feats, idx, n_units)
        ok = tuple(f_all.shape) == (n_units, 48, 28) and i_all.dtype == torch.int64
        for u in range(n_units):
            ok = ok and torch.equal(f_all[u], _patch_tensors(u)[0]) and torch.equal(i_all[u], _patch_tensors(u)[1])
        q.put((rank, ok, mine))
    finally:
        dist.destroy_process_group()


def _run(n_units, worker=None):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=worker or _worker, args=(r, 2, port, n_units, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    return sorted(res)


def test_shard_units_round_robin():
    assert cd.shard_units(32, 3, 8) == [3, 11, 19, 27]
    assert cd.shard_units(5, 1, 2) == [1, 3]
    assert cd.shard_units(1, 1, 2) == []
    assert sorted(sum((cd.shard_units(13, r, 4) for r in range(4)), [])) == list(range(13))


def test_all_gather_of_patch_records_even_and_ragged():
    for n_units in (4, 5):          # 5: rank 1 owns fewer patches than rank 0 (padding path)
        res = _run(n_units)
        assert [r[1] for r in res] == [True, True], res
        assert res[0][2] == list(range(0, n_units, 2)) and res[1][2] == list(range(1, n_units, 2))


def test_all_gather_of_patch_tensors_in_unit_order():
    """The tensor form of the exchange (what LocalSPFN.run_shape_sharded uses): even, ragged, and a rank without
    any patch."""
    for n_units in (4, 5, 1):
        res = _run(n_units, worker=_worker_tensors)
        assert [r[1] for r in res] == [True, True], (n_units, res)
