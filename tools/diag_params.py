"""Diagnostic (GPU box): distribution of the fitted-parameter differences at the bench configuration --
GPU pipeline vs oracle fitters on the same memberships (e_same), vs the all-oracle pipeline (e_all), and the
oracle's own sensitivity to the MLP-path difference of its inputs (sens).  Writes gpurun_out/diag_params.json."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpfn_b200 import api, synth  # noqa: E402
from oracle import fitters as ofit, network as onet  # noqa: E402
from tests.test_gpu_bench_config import _param_err  # noqa: E402

dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False
out_all = {}
for B, N, seed in ((16, 8192, 1234), (2, 1024, 51)):
    eng = api.GlobalSPFN(output_sizes=[3, 4, 28], device=dev)
    sd = {k: torch.from_numpy(v) for k, v in synth.network_state(eng.model.state_dict(), seed=1234 if B == 16 else 7).items()}
    eng.load_state_dict(sd)
    P = synth.shape_batch(B, N, seed=seed, k_slots=28)[0]
    out = eng.forward(torch.from_numpy(P).to(dev), dropout=False)
    ref = onet.pointnet2_forward(sd, P, 3)
    got = {k: v.cpu().numpy() for k, v in out["parameters"].items()}
    Xn, _, Wn = onet.spfn_postprocess(ref["heads"])
    same = ofit.compute_parameters(P, out["W"].cpu().numpy(), out["X"].cpu().numpy())
    allo = ofit.compute_parameters(P, Wn, Xn)
    sens, e_same, e_all = _param_err(same, allo), _param_err(got, same), _param_err(got, allo)
    q = [0.1, 0.5, 0.9, 1.0]
    res = {}
    for k in got:
        res[k] = {"sens": np.quantile(sens[k], q).tolist(), "e_same": np.quantile(e_same[k], q).tolist(),
                  "e_all": np.quantile(e_all[k], q).tolist(),
                  "e_same_where_sens<1e-4": float(e_same[k][sens[k] < 1e-4].max()) if (sens[k] < 1e-4).any() else None,
                  "frac_sens<1e-4": float((sens[k] < 1e-4).mean()), "frac_sens<1e-3": float((sens[k] < 1e-3).mean()),
                  "e_all_where_sens<1e-3": float(e_all[k][sens[k] < 1e-3].max()) if (sens[k] < 1e-3).any() else None}
    out_all["B%d_N%d" % (B, N)] = res
    print(json.dumps(res, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out_all, open("gpurun_out/diag_params.json", "w"), indent=1)
