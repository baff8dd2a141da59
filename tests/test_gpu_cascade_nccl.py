"""The patch-sharded LocalSPFN cascade (evaluation_localSPFN.py:95-130) under torch.distributed with the
NCCL backend: ``LocalSPFN.run_shape_sharded`` on G ranks (patches i mod G, one all-gather of the per-point outputs,
merge on rank 0; the exchange either as an NCCL all-gather or fused into the producing kernel as peer-to-peer writes
into the merge rank's symmetric-memory buffers) must return exactly what the single-GPU ``run_shape`` returns -- the backbone's per-patch outputs do
not depend on which other patches share the batch, so labels, fused memberships, normals and types are bit-identical.
G = 1 always runs (a world of one rank); G = 2 needs two visible GPUs (``gpurun --gpus 2``)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _shape(Ng=20000, nb=6, Kg=28, seed=3):
    from cpfn_b200 import synth
    P, Xn, _, I = synth.shape_batch(1, Ng, seed=seed)
    P, Xn, I = P[0].astype(np.float32), Xn[0].astype(np.float32), I[0]
    rng = np.random.RandomState(seed)
    seeds = P[rng.choice(Ng, nb, replace=False)]
    S = np.eye(Kg, dtype=np.int64)[I % Kg]
    types = rng.randn(Ng, 4).astype(np.float32)
    return P, seeds, S, Xn, types


def _worker(rank, world, port, nb, q):
    import torch.distributed as dist
    from cpfn_b200 import api, synth
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    import datetime
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev, timeout=datetime.timedelta(seconds=90))
    try:
        loc = api.LocalSPFN(n_max_local_instances=21, device=dev, num_points_patch=2048)
        sd = {k: torch.from_numpy(v) for k, v in synth.network_state(loc.engine.model.state_dict(), seed=11).items()}
        loc.load_state_dict(sd)
        P, seeds, S, Xn, types = _shape(nb=nb)
        t = lambda a: torch.from_numpy(a).to(dev)
        ok, why = True, ""
        for graphed, exchange in ((False, "nccl"), (True, "nccl"), (False, "auto"), (True, "auto"), (True, "auto")):
            res = loc.run_shape_sharded(t(P), t(S), t(Xn), t(types), seeds=t(seeds), dropout=False, graphed=graphed,
                                        exchange=exchange)
            if rank == 0:
                if world > 1 and res["exchange"] != ("nccl" if exchange == "nccl" else "p2p"):
                    ok, why = False, why + " exchange=%s ran as %s (%s)" % (exchange, res["exchange"],
                                                                          loc.__dict__.get("_exchange_error"))
                want = loc.run_shape(t(P), t(S), t(Xn), t(types), seeds=t(seeds), dropout=False)
                for k in ("W", "X", "T", "W_fusion", "X_global", "T_global"):
                    if res[k].shape != want[k].shape or not torch.equal(res[k], want[k]):
                        bad = (res[k] != want[k]).reshape(res[k].shape[0], -1).any(1).nonzero().flatten()[:8].tolist() \
                            if res[k].shape == want[k].shape else "shape %s vs %s" % (tuple(res[k].shape), tuple(want[k].shape))
                        ok, why = False, why + " %s(graphed=%s, %s, rows %s)" % (k, graphed, exchange, bad)
                if not np.array_equal(res["labels"], want["labels"]) or not torch.equal(
                        res["patch_indices"], want["patch_indices"].to(torch.int64)):
                    ok, why = False, why + " labels/indices(graphed=%s)" % graphed
            elif res is not None:
                ok, why = False, "non-merge rank returned a result"
        # patch indices handed in instead of seeds, merge on the last rank (every rank makes the same collective calls
        # whatever it found before)
        from cpfn_b200 import sampling_utils
        idx = sampling_utils.extract_patches(t(P), t(seeds), 2048)
        res = loc.run_shape_sharded(t(P), t(S), t(Xn), t(types), patch_indices=idx, dropout=False, merge_rank=world - 1)
        if rank == world - 1:
            want = loc.run_shape(t(P), t(S), t(Xn), t(types), patch_indices=idx, dropout=False)
            if not (torch.equal(res["W_fusion"], want["W_fusion"]) and np.array_equal(res["labels"], want["labels"])):
                ok, why = False, why + " patch_indices path"
        q.put((rank, ok, why))
    except Exception as e:                       # report instead of leaving the other rank in a collective
        import traceback
        q.put((rank, False, "%s\n%s" % (e, traceback.format_exc())))
    finally:
        try:
            dist.destroy_process_group()
        except Exception:
            pass


def _run(world, nb):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, nb, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = []
    try:
        for _ in ps:
            res.append(q.get(timeout=240))
    except Exception:
        res.append((-1, False, "timeout waiting for a rank"))
    for p in ps:
        p.join(timeout=30)
        if p.is_alive():
            p.kill()
    return sorted(res)


@pytest.mark.parametrize("world,nb", [(1, 5), (2, 6), (2, 5), (2, 1), (4, 6)])
def test_sharded_shape_equals_single_gpu(cuda_dev, world, nb):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    res = _run(world, nb)
    assert all(r[1] for r in res), res
