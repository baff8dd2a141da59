"""The data-parallel training step (cpfn_b200/train.py; Utils/training_utils.py:84-158 of the reference): the step
runs on this package's kernels with the reference's own loss glue, and gradients averaged over two NCCL ranks equal
the single-GPU gradients of the whole batch (BatchNorm frozen, so that sharding does not change the statistics)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _reference_root():
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for root in ("/root/reference", os.path.join(here, "baseline", "_ref")):
        if os.path.isfile(os.path.join(root, "SPFN", "losses_implementation.py")):
            return root
    return None


def _setup(dev, seed=5, K=21):
    from cpfn_b200 import synth
    from cpfn_b200.pn2_network import PointNet2
    root = _reference_root()
    if root not in sys.path:
        sys.path.insert(0, root)                       # the loss glue is the reference's own code
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    model = PointNet2(dim_input=3, dim_pos=3, output_sizes=[3, 4, K]).to(dev)
    sd = {k: torch.from_numpy(v) for k, v in synth.network_state(model.state_dict(), seed=seed).items()}
    model.load_state_dict(sd)
    return model


def _batch(dev, B=4, N=2048, K=21, seed=8):
    from cpfn_b200 import synth
    b = synth.training_batch(B, N, seed, k_slots=K, n_gt_points=128)
    t = lambda a: torch.from_numpy(a).to(dev)
    out = {k: t(v) for k, v in b.items() if k != "gt_parameters"}
    out["gt_parameters"] = {k: t(v) for k, v in b["gt_parameters"].items()}
    return out


def test_train_step_single_gpu(cuda_dev):
    if _reference_root() is None:
        pytest.skip("the loss glue needs the reference checkout (or baseline/_ref)")
    from cpfn_b200 import train
    model = _setup(cuda_dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    batch = _batch(cuda_dev)
    before = model.fc2[2].weight.detach().clone()
    losses = [float(train.train_step(model, opt, batch, train.LOCAL_MULTIPLIERS)[0].detach()) for _ in range(3)]
    assert all(np.isfinite(l) for l in losses) and losses[2] < losses[0]
    assert not torch.equal(before, model.fc2[2].weight)
    # the Global configuration also runs the differentiable fitters and the residue loss
    tm = {}
    out = train.train_step(model, opt, batch, train.GLOBAL_MULTIPLIERS, timings=tm)
    assert np.isfinite(float(out[0])) and float(out[4]) >= 0 and set(tm) == {"start", "forward", "backward", "all_reduce", "step"}


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import datetime
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev, timeout=datetime.timedelta(seconds=90))
    try:
        from cpfn_b200 import pn2_network, train
        # the always-on dropout draws a different mask for a different batch shape: identity for this comparison
        pn2_network.F = type("F", (), {"dropout": staticmethod(lambda x, p=0.5, **kw: x), "relu": staticmethod(torch.relu)})
        model = _setup(dev).eval()                       # BatchNorm frozen: sharding must not change the statistics
        batch = _batch(dev)
        mine = train.shard_batch(batch, rank, world)
        model.zero_grad()
        train.forward_losses(model, mine)[0].backward()
        train.all_reduce_gradients(model)
        ok, why = True, ""
        if rank == 0:
            ref = _setup(dev).eval()
            train.forward_losses(ref, batch)[0].backward()
            for (n, p), (_, r) in zip(model.named_parameters(), ref.named_parameters()):
                scale = float(r.grad.abs().max()) + 1e-12
                if float((p.grad - r.grad).abs().max()) > 2e-4 * scale + 1e-7:
                    ok, why = False, why + " " + n
        q.put((rank, ok, why))
    except Exception as e:
        import traceback
        q.put((rank, False, "%s\n%s" % (e, traceback.format_exc())))
    finally:
        try:
            dist.destroy_process_group()
        except Exception:
            pass


def test_sharded_gradients_equal_single_gpu(cuda_dev):
    if _reference_root() is None:
        pytest.skip("the loss glue needs the reference checkout (or baseline/_ref)")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = []
    try:
        for _ in ps:
            res.append(q.get(timeout=240))
    except Exception:
        res.append((-1, False, "timeout"))
    for p in ps:
        p.join(timeout=30)
        if p.is_alive():
            p.kill()
    assert all(r[1] for r in res), [(r[0], r[1], str(r[2])[:600]) for r in res]
