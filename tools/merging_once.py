"""One pass of the merge kernels at the LocalSPFN size (for the ncu launch list in profiles/r1_merging.md):
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:merge_ --csv python tools/merging_once.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cpfn_b200 import merging_utils, synth
dev = torch.device("cuda:0")
Ng, nb, Np, Kl, Kg = 131072, 32, 8192, 21, 28
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
for _ in range(2):
    merging_utils.merge_shape(W, X, T, idx.to(torch.int32), S, on, ot)
torch.cuda.synchronize()
