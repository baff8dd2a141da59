"""Wall-clock of run_host variants (experiment; not part of the bench)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cpfn_b200 import api, synth

dev = torch.device("cuda:0")
B, N, steps = 16, 8192, 200
host = [torch.from_numpy(synth.clouds(B, N, seed=s)).pin_memory() for s in range(4)] if hasattr(synth, "clouds") else \
       [torch.from_numpy(synth.shape_batch(B, N, seed=s)[0]).pin_memory() for s in range(4)]

def timeit(name, fn):
    for i in range(5):
        fn(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        fn(i)
    torch.cuda.synchronize()
    print("%-40s %.1f us/step" % (name, (time.perf_counter() - t0) / steps * 1e6), flush=True)

eng = api.GlobalSPFN(device="cuda:0")
Pd = host[0].to(dev)
def replay_only(i):
    eng.forward_graphed(Pd); torch.cuda.synchronize()
timeit("graph replay + D2D + sync", replay_only)
def v0(i):
    P = host[i % 4].to(dev, non_blocking=True)
    out = eng.forward_graphed(P)
    a = eng._pinned("a", out["X"].shape, torch.float32); a.copy_(out["X"], non_blocking=True)
    b = eng._pinned("b", out["instance"].shape, torch.int32); b.copy_(out["instance"], non_blocking=True)
    c = eng._pinned("c", out["type"].shape, torch.int32); c.copy_(out["type"], non_blocking=True)
    d = eng._pinned("d", out["parameters_packed"].shape, torch.float32); d.copy_(out["parameters_packed"], non_blocking=True)
    torch.cuda.synchronize()
timeit("v0 old (stage.to + D2D, D2H after)", v0)
timeit("v1 direct H2D, D2H after graph", lambda i: eng.run_host(host[i % 4], graphed=True, overlap_d2h=False))
timeit("v2 two graphs, D2H overlaps the fitters", lambda i: eng.run_host(host[i % 4], graphed=True))
def h2d_only(i):
    Pd.copy_(host[i % 4], non_blocking=True); torch.cuda.synchronize()
timeit("H2D 1.5 MB + sync", h2d_only)
X = torch.empty(B, N, 3, device=dev); hx = torch.empty(B, N, 3).pin_memory()
def d2h_only(i):
    hx.copy_(X, non_blocking=True); torch.cuda.synchronize()
timeit("D2H 1.5 MB + sync", d2h_only)
