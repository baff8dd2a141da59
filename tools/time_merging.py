"""Device time of the merge kernels at the LocalSPFN size (CUDA events, L2 flushed between repetitions) with the
reference's dense formulation timed beside them: torch fp32 on the same GPU (the 'before') and numpy on the host."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from cpfn_b200 import merging_utils, synth

dev = torch.device("cuda:0")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts))


rows = []
for Ng in (131072, 1 << 20):
    nb, Np, Kl, Kg = 32, 8192, 21, 28
    g = torch.Generator(device="cpu").manual_seed(5)
    P = torch.from_numpy(synth.shape_cloud(Ng, 99)[0]).to(dev)
    idx = torch.empty(nb, Np, dtype=torch.int64, device=dev)
    for b in range(nb):
        c = P[int(torch.randint(Ng, (1,), generator=g))]
        idx[b] = torch.topk((P - c).norm(dim=1), Np, largest=False).indices
    W = torch.softmax(4 * torch.randn(nb, Np, Kl, generator=g).to(dev), dim=2)
    S = torch.nn.functional.one_hot(torch.randint(Kg, (Ng,), generator=g), Kg).float().to(dev)
    X = torch.nn.functional.normalize(torch.randn(nb, Np, 3, generator=g), dim=2).to(dev)
    T = torch.randn(nb, Np, 4, generator=g).to(dev)
    on = torch.nn.functional.normalize(torch.randn(Ng, 3, generator=g), dim=1).to(dev)
    ot = torch.randn(Ng, 4, generator=g).to(dev)
    idx32 = idx.to(torch.int32)
    inverse = merging_utils.inverse_index(idx32, Ng)
    sim = merging_utils.similarity_soft(S, W, idx32, inverse=inverse)
    t0 = time.perf_counter()
    labels = merging_utils.run_heuristic_solver(sim.cpu().numpy(), nb, Kg, Kl)
    solver_ms = (time.perf_counter() - t0) * 1e3
    r = {"N_global": Ng, "patches": nb, "labels_out": int(labels.max()) + 1,
         "inverse_index_us": timed(lambda: merging_utils.inverse_index(idx32, Ng)),
         "similarity_us": timed(lambda: merging_utils.similarity_soft(S, W, idx32, inverse=inverse)),
         "fuse_patches_us": timed(lambda: merging_utils.fuse_patches(S, W, idx32, labels, inverse=inverse)),
         "normals_types_us": timed(lambda: merging_utils.merge_normals_types(X, T, idx32, on, ot, inverse=inverse)),
         "host_solver_ms": round(solver_ms, 2)}
    M = nb * Kl + Kg

    def dense_similarity():                      # merging_utils.py:6-15 as written, torch fp32 on this GPU
        A = torch.zeros(Ng, M, device=dev)
        for b in range(nb):
            A[idx[b], b * Kl:(b + 1) * Kl] += W[b]
        A[:, nb * Kl:] = S
        return torch.mm(A.transpose(0, 1), A)
    if Ng <= 131072 or torch.cuda.mem_get_info()[0] > 8 << 30:
        r["reference_formulation_torch_gpu_similarity_us"] = timed(dense_similarity, reps=5)
    if Ng <= 131072:
        Wn, Sn, In = W.cpu().numpy(), S.cpu().numpy(), idx.cpu().numpy()
        t0 = time.perf_counter()
        A = np.zeros((Ng, M), np.float32)
        for b in range(nb):
            A[In[b], b * Kl:(b + 1) * Kl] += Wn[b]
        A[:, nb * Kl:] = Sn
        A.T @ A
        r["reference_formulation_numpy_host_similarity_ms"] = round((time.perf_counter() - t0) * 1e3, 1)
    # algorithmic bytes: W + indices + S read once, inverse index read per patch pair lookup
    r["similarity_algorithmic_bytes"] = nb * Np * Kl * 4 + nb * Np * 4 + Ng * Kg * 4
    r["similarity_GBps"] = round(r["similarity_algorithmic_bytes"] / r["similarity_us"] / 1e3, 1)
    rows.append({k: (round(v, 1) if isinstance(v, float) else v) for k, v in r.items()})
    print(rows[-1], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/merging_timing.json", "w"), indent=1)
