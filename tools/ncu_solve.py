import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpfn_b200 import api, merging_utils, sampling_utils, synth
dev = torch.device("cuda:0")
Ng = 131072
P, Xn, I = synth.shape_cloud(Ng, seed=4242)[:3]
rng = np.random.RandomState(7)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
seeds, Pg = t(P[rng.choice(Ng, 32, replace=False)]), t(P)
S = torch.nn.functional.one_hot(t(I % 28), 28).float()
idx = sampling_utils.extract_patches(Pg, seeds, 8192)
W = torch.softmax(torch.randn(32, 8192, 21, device=dev) * 3, dim=2)
if len(sys.argv) > 1:   # realistic memberships from the network
    loc = api.LocalSPFN(n_max_local_instances=21, device=dev)
    loc.load_state_dict({k: torch.from_numpy(v) for k, v in synth.network_state(loc.engine.model.state_dict(), seed=1234).items()})
    W = loc.engine.forward(api.LocalSPFN.normalise_patches(Pg, idx), dropout=True, fit=False)["W"]
inv = merging_utils.inverse_index(idx, Ng)
sim = merging_utils.similarity_soft(S, W, idx, inverse=inv)
torch.cuda.synchronize()
for _ in range(2):
    sol = merging_utils.solve_labels_device(sim, 32, 28, 21)
torch.cuda.synchronize()
print("n_labels", int(sol[2].item()), "pairs>0", int(((sim > 0).sum() - (sim.diagonal() > 0).sum()) // 2))
