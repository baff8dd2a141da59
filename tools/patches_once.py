"""One call of extract_patches per size (for ncu launch lists)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cpfn_b200 import sampling_utils, synth
dev = torch.device("cuda:0")
for N, S in ((131072, 1), (131072, 32), (1 << 20, 1), (1 << 20, 32)):
    hr_d = torch.from_numpy(synth.shape_cloud(N, 77)[0].astype(np.float32)).to(dev)
    seeds_d = hr_d[:: N // S][:S].contiguous()
    sampling_utils.extract_patches(hr_d, seeds_d, 8192, return_distances=True)
    torch.cuda.synchronize()
