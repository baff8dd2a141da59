"""three_nn_weights at the FP3 shape (B = 16, n = 8192 queries, m = 512 known): grid kernel on the original query
order, the same on the ball-query grid's cell order (cpfn_three_nn_weights_sorted), and the exhaustive scan; CUDA
events, L2 flushed, outputs compared bit for bit."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpfn_b200 import _lib, cuda_ops, fused, synth
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
B, n, m = 16, 8192, 512
P = torch.from_numpy(synth.shape_batch(B, n, seed=1234)[0]).to(dev)
known = fused.gather_xyz(P, cuda_ops.farthest_point_sampling(P, m))
L = _lib.lib()
nbytes = L.cpfn_ball_query_grid_workspace_bytes(B, n)
ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
_lib.check(L.cpfn_ball_query_grid_build(P.data_ptr(), B, n, 0.2, ws.data_ptr(), nbytes, torch.cuda.current_stream(dev).cuda_stream), "build")
ref = fused.three_nn_weights(P, known)
for name, fn in (("grid, original order", lambda: fused.three_nn_weights(P, known)),
                 ("grid, cell-sorted queries", lambda: fused.three_nn_weights(P, known, sorted_queries=ws)),
                 ("exhaustive scan", None)):
    if fn is None:
        os.environ["CPFN_NN_NO_GRID"] = "1"
        fn = lambda: fused.three_nn_weights(P, known)
    out = fn(); torch.cuda.synchronize()
    same = bool(torch.equal(out[0], ref[0]) and torch.equal(out[1], ref[1]))
    ts = []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    print("%-28s %7.1f us   identical: %s" % (name, float(np.median(ts)), same), flush=True)
