"""Stage / kernel timing of the LocalSPFN shape cascade on one GPU (CUDA events, L2 not flushed: the stages follow
each other on hot data exactly as in run_shape).  Writes gpurun_out/cascade_timing.json."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpfn_b200 import api, merging_utils, sampling_utils, synth  # noqa: E402

dev = torch.device("cuda:0")
out = {}


def timed(fn, n=20, warm=3):
    for _ in range(warm):
        r = fn()
    torch.cuda.synchronize()
    ev = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); r = fn(); b.record()
        ev.append((a, b))
    torch.cuda.synchronize()
    return r, float(np.median([a.elapsed_time(b) for a, b in ev])) * 1e3


for Ng in (131072, 1 << 20):
    loc = api.LocalSPFN(n_max_local_instances=21, device=dev)
    sd = {k: torch.from_numpy(v) for k, v in synth.network_state(loc.engine.model.state_dict(), seed=1234).items()}
    loc.load_state_dict(sd)
    P, Xn, I = synth.shape_cloud(Ng, seed=4242)[:3]
    rng = np.random.RandomState(7)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    seeds, Pg, Xg = t(P[rng.choice(Ng, 32, replace=False)]), t(P), t(Xn)
    S = torch.nn.functional.one_hot(t(I % 28), 28).float()
    Tg = t(rng.randn(Ng, 4).astype(np.float32))
    r = {}
    idx, r["extract_patches"] = timed(lambda: sampling_utils.extract_patches(Pg, seeds, 8192))
    Pn, r["normalise_patches"] = timed(lambda: api.LocalSPFN.normalise_patches(Pg, idx))
    net, r["backbone_graphed_B32"] = timed(lambda: loc.engine.forward_graphed(Pn, dropout=True, fit=False))
    W, X, T = net["W"].clone(), net["X"].clone(), net["T"].clone()
    inv, r["inverse_index"] = timed(lambda: merging_utils.inverse_index(idx, Ng))
    sim, r["similarity"] = timed(lambda: merging_utils.similarity_soft(S, W, idx, inverse=inv))
    sol, r["solve_labels_device"] = timed(lambda: merging_utils.solve_labels_device(sim, 32, 28, 21))
    _, r["normals_types"] = timed(lambda: merging_utils.merge_normals_types(X, T, idx, Xg, Tg, inverse=inv))
    L, r["n_labels_item_sync"] = timed(lambda: int(sol[2].item()))
    _, r["fuse_patches"] = timed(lambda: merging_utils.fuse_patches(S, W, idx, None, inverse=inv, device_solution=(sol[0], sol[1], L)))
    _, r["host_solver_incl_d2h"] = timed(lambda: merging_utils.run_heuristic_solver(sim.cpu().numpy(), 32, 28, 21), n=5)
    _, r["merge_shape_device"] = timed(lambda: merging_utils.merge_shape(W, X, T, idx, S, Xg, Tg))
    _, r["merge_shape_host"] = timed(lambda: merging_utils.merge_shape(W, X, T, idx, S, Xg, Tg, solver="host"), n=5)
    _, r["run_shape_total"] = timed(lambda: loc.run_shape_sharded(Pg, S, Xg, Tg, seeds=seeds, dropout=True, graphed=True))
    r["n_labels"] = L
    out["Ng_%d" % Ng] = r
    print(Ng, json.dumps(r, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/cascade_timing.json", "w"), indent=1)
