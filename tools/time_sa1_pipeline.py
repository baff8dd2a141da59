"""Where the pipelined SA1 spends its time: sampling alone (1 launch / n launches), ball query + chain alone, and the
pipelined whole, CUDA events around the region on the main stream (all streams joined)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpfn_b200 import _lib, cuda_ops, fused, synth
from cpfn_b200.pn2_network import PointNet2
dev = torch.device("cuda:0")
L = _lib.lib()
model = PointNet2(output_sizes=[3, 4, 28]).to(dev).eval()
model.load_state_dict({k: torch.from_numpy(v) for k, v in synth.network_state(model.state_dict(), seed=4).items()})
P = torch.from_numpy(synth.shape_batch(16, 8192, seed=1234)[0]).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
main = torch.cuda.current_stream(dev)

def timed(fn, n=10):
    ts = []
    for _ in range(n + 2):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts[2:]))

B, N, S = 16, 8192, 512
idx = torch.empty(B, S, dtype=torch.int32, device=dev); xyz = torch.empty(B, S, 3, device=dev); state = torch.empty(B, N, device=dev)
def fps_chunks(n, floor):
    per = S // n
    for c in range(n):
        _lib.check(L.cpfn_furthest_point_sampling_rounds(P.data_ptr(), B, N, S, c * per, (c + 1) * per, idx.data_ptr(), xyz.data_ptr(),
                                                         state.data_ptr(), state.numel() * 4, floor, main.cuda_stream), "r")
with torch.no_grad():
    print("fps one launch (old entry)        %.1f us" % timed(lambda: cuda_ops.farthest_point_sampling(P, 512, return_centroids=True)))
    for n in (1, 2, 4, 8):
        print("fps rounds x%d floor 0              %.1f us" % (n, timed(lambda: fps_chunks(n, 0))))
    print("fps rounds x4 floor 120K           %.1f us" % timed(lambda: fps_chunks(4, 120 * 1024)))
    ind = fused.sa_indices(model.sa1, P)
    print("sa1 ball query + chain (all SMs)   %.1f us" % timed(lambda: fused.sa_forward_pm(model.sa1, P, None, indices=(ind[0], cuda_ops.ball_query(ind[0], P, 0.2, 64)))))
    print("sa1 chain only                     %.1f us" % timed(lambda: fused.sa_forward_pm(model.sa1, P, None, indices=ind)))
    w = torch.cuda.Stream(device=dev)
    def piped(n):
        r = fused.sa1_pipelined(model.sa1, P, w, n)
        main.wait_event(r[2])
    for n in (2, 4, 8):
        print("sa1 pipelined x%d                   %.1f us" % (n, timed(lambda: piped(n))))
    for fl in ("0", "64", "100", "160"):
        os.environ["CPFN_FPS_FLOOR_KB"] = fl
        print("sa1 pipelined x4 floor %sK         %.1f us" % (fl, timed(lambda: piped(4))))

# ---- per-kernel timeline of the pipelined form (events on both streams, ms relative to the start) ----
print("timeline (us from start): kernel begin-end")
K, cout = 64, 128
pc = fused._sa_chain(model.sa1, dev)
n_chunks, per = 4, 128
for trial in range(3):
    fps_idx = torch.empty(B, S, dtype=torch.int32, device=dev); new_xyz = torch.empty(B, S, 3, device=dev)
    state = torch.empty(B, N, device=dev); gidx = torch.empty(B, S, K, dtype=torch.int32, device=dev)
    out = torch.zeros(B, S, cout, device=dev)
    nbytes = L.cpfn_ball_query_grid_workspace_bytes(B, N); ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    _lib.check(L.cpfn_ball_query_grid_build(P.data_ptr(), B, N, 0.2, ws.data_ptr(), nbytes, main.cuda_stream), "b")
    flush.zero_(); torch.cuda.synchronize()
    evs = []
    def ev(stream):
        e = torch.cuda.Event(enable_timing=True); e.record(stream); return e
    t0 = ev(main)
    for c in range(n_chunks):
        j0, j1 = c * per, (c + 1) * per
        a = ev(main)
        _lib.check(L.cpfn_furthest_point_sampling_rounds(P.data_ptr(), B, N, S, j0, j1, fps_idx.data_ptr(), new_xyz.data_ptr(),
                                                         state.data_ptr(), state.numel() * 4, 120 * 1024, main.cuda_stream), "r")
        b_ = ev(main)
        evs.append(("fps%d" % c, a, b_))
        with torch.cuda.stream(w):
            w.wait_event(b_)
            a2 = ev(w)
            _lib.check(L.cpfn_ball_query_grid_query_range(new_xyz.data_ptr(), P.data_ptr(), B, N, S, j0, per, 0.2, K, gidx.data_ptr(),
                                                          ws.data_ptr(), nbytes, w.cuda_stream), "q")
            b2 = ev(w)
            fused.run_chain(pc, B, S * K, out, cout, tile_cols=128, in_mode=fused.IN_GROUP, a_src=None, a_ch=0, a_rows=N, idx=gidx,
                            xyz=P, centers=new_xyz, group_k=K, out_mode=fused.OUT_POOL, pool_g=K, l0=pc.l0, out_prezeroed=True,
                            window=(j0 * K, per * K), max_ctas=(168 if c + 1 < n_chunks else 0))
            c2 = ev(w)
            evs.append(("bq%d" % c, a2, b2)); evs.append(("chain%d" % c, b2, c2))
    torch.cuda.synchronize()
    if trial == 2:
        for name, a, b_ in evs:
            print("  %-7s %7.1f - %7.1f" % (name, t0.elapsed_time(a) * 1e3, t0.elapsed_time(b_) * 1e3))
