"""FPS variants at the SA1 shape (B clouds x 8192 points -> 512 samples): CUDA-event times, L2 flushed, bit-exact
check against the single-cloud cluster kernel.  Writes gpurun_out/fps_variants.json."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpfn_b200 import cuda_ops, synth
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = {}
for B in (16, 32):
    P = torch.from_numpy(synth.shape_batch(B, 8192, seed=1234)[0]).to(dev)
    os.environ["CPFN_FPS_PAIR"] = "0"; os.environ["CPFN_FPS_POLL"] = "0"; os.environ.pop("CPFN_FPS_CLUSTER", None)
    ref = cuda_ops.farthest_point_sampling(P, 512).clone()
    for name, env in (("single_mbar_C4", {"CPFN_FPS_POLL": "0"}), ("single_poll_C4", {}), ("single_poll_C8", {"CPFN_FPS_CLUSTER": "8"}),
                      ("single_poll_C2", {"CPFN_FPS_CLUSTER": "2"}),
                      ("pair_poll_C8", {"CPFN_FPS_PAIR": "1", "CPFN_FPS_CLUSTER": "8"}),
                      ("pair_poll_C4", {"CPFN_FPS_PAIR": "1", "CPFN_FPS_CLUSTER": "4"})):
        for k in ("CPFN_FPS_CLUSTER", "CPFN_FPS_PAIR", "CPFN_FPS_POLL"):
            os.environ.pop(k, None)
        os.environ.update(env)
        try:
            got = cuda_ops.farthest_point_sampling(P, 512)
            torch.cuda.synchronize()
        except Exception as e:
            out["B%d_%s" % (B, name)] = str(e); continue
        ts = []
        for _ in range(10):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); cuda_ops.farthest_point_sampling(P, 512); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        out["B%d_%s" % (B, name)] = {"us": float(np.median(ts)), "exact": bool(torch.equal(got, ref))}
        print(B, name, out["B%d_%s" % (B, name)], flush=True)
json.dump(out, open("gpurun_out/fps_variants.json", "w"), indent=1)
