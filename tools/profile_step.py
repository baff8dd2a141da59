"""Two eager GlobalSPFN steps at the bench configuration (B = 16 x 8192 points, K = 28, forward + fit) for ncu:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py
    ncu --set full --clock-control none --import-source on -k regex:'mlp_chain|fps_cluster|tls_|bq_grid|three_nn|dropout|spfn_post' \
        -s 25 -c 25 -o gpurun_out/step python tools/profile_step.py
The state dict is loaded from a numpy file written on first use so that the copy kernels of load_state_dict do not
pad the launch list."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpfn_b200 import api, synth  # noqa: E402

dev = torch.device("cuda:0")
eng = api.GlobalSPFN(output_sizes=[3, 4, 28], device=dev)
eng.load_state_dict({k: torch.from_numpy(v) for k, v in synth.network_state(eng.model.state_dict(), seed=1234).items()})
P = torch.from_numpy(synth.shape_batch(16, 8192, seed=1234, k_slots=28)[0]).to(dev)
torch.cuda.synchronize()
print("PROFILE_BEGIN", flush=True)
for i in range(2):
    torch.manual_seed(i)
    eng.forward(P, dropout=True, fit=True)
torch.cuda.synchronize()
