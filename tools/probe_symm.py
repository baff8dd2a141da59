"""Probe (2 GPUs, torchrun): does torch's symmetric memory give peer-mapped buffers on this box, and can one of this
library's kernels write straight into the peer's HBM?"""
import datetime
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=60))
import torch.distributed._symmetric_memory as symm
try:
    n = 1 << 20
    t = symm.empty(n, dtype=torch.float32, device=dev)
    hdl = symm.rendezvous(t, dist.group.WORLD)
    print(rank, "rendezvous ok", type(hdl).__name__, [hex(p) for p in hdl.buffer_ptrs][:4], flush=True)
    t.fill_(float(rank))
    hdl.barrier(channel=0)
    peer = hdl.get_buffer((rank + 1) % world, (n,), torch.float32)
    print(rank, "peer value", float(peer[0]), "ptr", hex(peer.data_ptr()), flush=True)
    # one of our kernels writing into the peer buffer: linear_rows (out = bias + x W^T)
    from cpfn_b200 import fused
    x = torch.ones(4, 8, device=dev)
    W = torch.ones(16, 8, device=dev)
    b = torch.full((16,), float(rank), device=dev)
    out_peer = peer[:4 * 16].view(4, 16)
    fused.linear_rows(x, W, b, out_peer)
    hdl.barrier(channel=1)
    print(rank, "mine after peer wrote", t[:3].tolist(), "(expect %g)" % (8.0 + (rank - 1) % world), flush=True)
    # timing: push 32 MB to the peer with a copy kernel + barrier
    src = torch.randn(8 << 20, device=dev)
    big = symm.empty(8 << 20, dtype=torch.float32, device=dev)
    h2 = symm.rendezvous(big, dist.group.WORLD)
    pb = h2.get_buffer((rank + 1) % world, (8 << 20,), torch.float32)
    for _ in range(3):
        pb.copy_(src); h2.barrier(channel=0)
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        pb.copy_(src); h2.barrier(channel=0)
    e.record(); torch.cuda.synchronize()
    print(rank, "32 MB peer copy + barrier: %.1f us" % (a.elapsed_time(e) * 100), flush=True)
    recv = torch.empty(world * (8 << 20), device=dev)
    for _ in range(3):
        dist.all_gather_into_tensor(recv, src)
    torch.cuda.synchronize()
    a.record()
    for _ in range(10):
        dist.all_gather_into_tensor(recv, src)
    e.record(); torch.cuda.synchronize()
    print(rank, "NCCL all_gather of 32 MB per rank: %.1f us" % (a.elapsed_time(e) * 100), flush=True)
except Exception as ex:
    import traceback
    print(rank, "SYMM FAILED:", ex, traceback.format_exc(), flush=True)
dist.destroy_process_group()
