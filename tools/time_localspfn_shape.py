"""One shape through the LocalSPFN path (api.LocalSPFN.run_shape: seeds -> patches -> normalise -> backbone on 32
patches -> merge) at the sizes of BASELINE configs 3 / 5, stage by stage (CUDA events; the host-side greedy
label merge in wall clock)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from cpfn_b200 import api, merging_utils, sampling_utils, synth

dev = torch.device("cuda:0")
rows = []
loc = api.LocalSPFN(n_max_local_instances=21, device=dev)
for Ng in (131072, 1 << 20):
    P, Xn, _, I = synth.shape_batch(1, Ng, seed=8)
    P, Xn, I = P[0].astype(np.float32), Xn[0].astype(np.float32), I[0]
    rng = np.random.RandomState(3)
    t = lambda a: torch.from_numpy(a).to(dev)
    Pd, S, on, ot = t(P), t(np.eye(28, dtype=np.float32)[I % 28]), t(Xn), t(rng.randn(Ng, 4).astype(np.float32))
    seeds = Pd[torch.from_numpy(rng.choice(Ng, 32, replace=False)).to(dev)].contiguous()

    def ev():
        e = torch.cuda.Event(enable_timing=True); e.record(); return e
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0 = ev()
        idx = sampling_utils.extract_patches(Pd, seeds, 8192)
        e1 = ev()
        Pn = api.LocalSPFN.normalise_patches(Pd, idx)
        e2 = ev()
        out = loc.engine.forward_graphed(Pn, dropout=True, fit=False)
        e3 = ev()
        inverse = merging_utils.inverse_index(idx, Ng)
        sim = merging_utils.similarity_soft(S, out["W"], idx, inverse=inverse)
        e4 = ev()
        sim_host = sim.cpu().numpy()
        th = time.perf_counter()
        labels = merging_utils.run_heuristic_solver(sim_host, 32, 28, 21)
        solver_ms = (time.perf_counter() - th) * 1e3
        e5 = ev()
        Wf = merging_utils.fuse_patches(S, out["W"], idx, labels, inverse=inverse)
        Xg, Tg = merging_utils.merge_normals_types(out["X"], out["T"], idx, on, ot, inverse=inverse)
        e6 = ev()
        torch.cuda.synchronize()
        wall_ms = (time.perf_counter() - t0) * 1e3
    rows.append({"N_global": Ng, "patches": 32, "extract_patches_us": round(e0.elapsed_time(e1) * 1e3, 1),
                 "normalise_us": round(e1.elapsed_time(e2) * 1e3, 1), "backbone_32x8192_us": round(e2.elapsed_time(e3) * 1e3, 1),
                 "inverse_plus_similarity_us": round(e3.elapsed_time(e4) * 1e3, 1), "host_label_merge_ms": round(solver_ms, 2),
                 "fuse_labels_normals_types_us": round(e5.elapsed_time(e6) * 1e3, 1), "wall_ms_whole_shape": round(wall_ms, 2),
                 "merged_labels": int(labels.max()) + 1})
    print(rows[-1], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/localspfn_shape_timing.json", "w"), indent=1)
