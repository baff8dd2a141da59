"""Benchmark of the CPFN hot path on B200: GlobalSPFN forward + primitive fitting.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one pass of the hot path over one batch: PointNet2 forward (FPS, ball query,
SA / FP layers, heads; always-on dropout as in the reference), softmax / normalise, and the
four primitive fitters, on B=16 synthetic 8192-point clouds per GPU (BASELINE.json configs[1]).
Metric: points per second (whole job: all ranks' points / max-over-ranks device time).
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_PER_GPU, N_POINTS, K_SLOTS = 16, 8192, 28
METRIC, UNIT = "globalspfn_forward_points_per_s", "points/s"
WORKLOAD = "GlobalSPFN forward + 4-type primitive fit, 8192-pt clouds, batch 16 per GPU (BASELINE configs[1])"


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 6:
                    self.rows.append(parts)
            except Exception:
                pass
            self._stop_evt.wait(0.1)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def make_inputs(n_batches, batch, seed):
    from cpfn_b200 import synth
    return [synth.shape_batch(batch, N_POINTS, seed=seed + i, k_slots=K_SLOTS)[0] for i in range(n_batches)]


def model_state(template):
    from cpfn_b200 import synth
    return {k: torch.from_numpy(v) for k, v in synth.network_state(template, seed=1234).items()}


# ----------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle port on the host cores (the only place bench.py runs
# anything under oracle/).
# ----------------------------------------------------------------------------------------------

def cpu_step(sd, P):
    from oracle import fitters as ofit
    from oracle import network as onet
    ref = onet.pointnet2_forward(sd, P, 3)
    Xn, _, Wn = onet.spfn_postprocess(ref["heads"])
    return ofit.compute_parameters(P, Wn, Xn)


def cpu_baseline(steps, warmup, batch):
    from cpfn_b200.pn2_network import PointNet2
    from oracle import index_ops
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    index_ops.set_threads(cores)
    sd = model_state(PointNet2(output_sizes=[3, 4, K_SLOTS]).state_dict())
    inputs = make_inputs(2, batch, seed=4321)
    for i in range(warmup):
        cpu_step(sd, inputs[i % 2])
    t0 = time.perf_counter()
    for i in range(steps):
        cpu_step(sd, inputs[i % 2])
    dt = (time.perf_counter() - t0) / max(1, steps)
    return {"value": batch * N_POINTS / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d steps of %d clouds x %d points (oracle/network.py torch-CPU MLPs + OpenMP C index ops, "
                      "oracle/fitters.py numpy fitters)" % (steps, batch, N_POINTS),
            "ms_per_step": dt * 1e3}


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    batch = 2
    cb = cpu_baseline(args.steps, args.warmup, batch)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample_batch": batch, "n_points": N_POINTS, "k_slots": K_SLOTS},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(line)


# ----------------------------------------------------------------------------------------------
# Our arm
# ----------------------------------------------------------------------------------------------

class OpTimer:
    """Per-op CUDA-event timing on the launching stream (used in a separate profiling pass
    after the timed region, so it does not perturb the headline number)."""

    def __init__(self):
        self.events = {}

    def wrap(self, mod, name, label=None):
        fn = getattr(mod, name)
        label = label or name

        def timed(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*a, **k)
            e1.record()
            self.events.setdefault(label, []).append((e0, e1))
            return r
        setattr(mod, name, timed)
        return fn

    def summary(self):
        torch.cuda.synchronize()
        return {k: [a.elapsed_time(b) * 1e3 for a, b in v] for k, v in self.events.items()}


def run_ours(args):
    rank, local_rank, world = dist_env()
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from cpfn_b200 import api, cuda_ops, fused
    from cpfn_b200.spfn import fit as fitmod

    eng = api.GlobalSPFN(output_sizes=[3, 4, K_SLOTS], device=dev)
    eng.load_state_dict(model_state(eng.model.state_dict()))
    n_in = 4
    host_inputs = [torch.from_numpy(p).pin_memory() for p in make_inputs(n_in, B_PER_GPU, seed=1234 + 100 * rank)]
    dev_inputs = [p.to(dev) for p in host_inputs]
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    use_graph = not os.environ.get("CPFN_BENCH_NO_GRAPH")

    def step(i, graphed=None):
        torch.manual_seed(1000 + i)
        if use_graph if graphed is None else graphed:
            return eng.forward_graphed(dev_inputs[i % n_in])
        return eng.forward(dev_inputs[i % n_in])

    sampler = ClockSampler(local_rank)      # samples clocks / throttle reasons from warm-up to the end of the GPU work
    if rank == 0:
        sampler.start()
    for i in range(max(3, args.warmup)):
        step(i)
    barrier()
    launches0 = cuda_ops.LAUNCHES
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.zero_()                       # L2 flush between timed iterations (outside the event pair)
        ev[i][0].record()
        step(i)
        ev[i][1].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = cuda_ops.LAUNCHES - launches0
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    per_op, step_us, clocks = {}, float("nan"), None
    if rank == 0:
        # ---- profiling pass (rank 0): per-op device times, dominant kernel, roofline ----
        timer = OpTimer()
        originals = [(cuda_ops, n, timer.wrap(cuda_ops, n)) for n in
                     ("farthest_point_sampling", "ball_query", "three_nn", "three_weighted_sum", "group_points", "gather_points")]
        originals.append((fitmod, "fit_primitives_packed", timer.wrap(fitmod, "fit_primitives_packed", "fit_primitives")))
        if fused.available():
            originals.append((fused, "run_chain", timer.wrap(fused, "run_chain", "mlp_chain")))
        fused.USE_SIDE_STREAM = False          # serialise the step so that per-op event times are meaningful
        tot = []
        for i in range(args.steps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            step(i, graphed=False)
            b.record()
            tot.append((a, b))
        per_op = timer.summary()
        fused.USE_SIDE_STREAM = True
        clocks = sampler.stop()
        step_us = float(np.mean([a.elapsed_time(b) * 1e3 for a, b in tot]))
        for mod, n, fn in originals:
            setattr(mod, n, fn)
        breakdown = {k: round(float(np.sum(v)) / args.steps, 2) for k, v in per_op.items()}
        breakdown["step_total_serialised_ungraphed"] = round(step_us, 2)
    barrier()
    # end-to-end through the public host-buffer API (H2D + forward + fit + D2H every step).  Two numbers:
    # the latency of one synchronous call (run_host) and the throughput of the streaming call (stream_host:
    # the same work per step with two batches in flight, the next batch's H2D under the current batch's compute)
    for i in range(max(5, args.warmup)):     # first replays of a fresh graph include its upload
        eng.run_host(host_inputs[i % n_in], graphed=use_graph)
    barrier()
    torch.manual_seed(2000)
    t0 = time.perf_counter()
    h2d = d2h = 0
    for i in range(args.steps):
        _, h2d, d2h = eng.run_host(host_inputs[i % n_in], graphed=use_graph)
    barrier()
    lat_s = time.perf_counter() - t0
    e2e_s, pipelined = lat_s, False
    if use_graph:
        for _ in eng.stream_host(host_inputs[i % n_in] for i in range(max(5, args.warmup))):
            pass
        barrier()
        t0 = time.perf_counter()
        n_done = 0
        for _, h2d, d2h in eng.stream_host(host_inputs[i % n_in] for i in range(args.steps)):
            n_done += 1
        barrier()
        e2e_s, pipelined = time.perf_counter() - t0, True
        assert n_done == args.steps

    t = torch.tensor([dev_ms, e2e_s * 1e3, lat_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, lat_ms = float(t[0]), float(t[1]), float(t[2])

    if rank != 0:
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
            dist.destroy_process_group()
        return

    # Dominant kernel of ours: SA1 furthest point sampling (one launch per step at N=8192, m=512).
    fps_us = [t_ for t_ in per_op.get("farthest_point_sampling", [])]
    fps_big = fps_us[0::2] if len(fps_us) >= 2 else fps_us       # calls alternate SA1 (8192->512), SA2 (512->128)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak, peak_src = 6650.0, "fallback"
    if os.path.exists(peaks_path):
        with open(peaks_path) as f:
            peak, peak_src = float(json.load(f)["hbm_gbs"]), "measured"
    fps_bytes = B_PER_GPU * (512 - 1) * N_POINTS * 16
    fps_t = float(np.mean(fps_big)) * 1e-6 if fps_big else float("nan")
    achieved = fps_bytes / fps_t / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r1_traffic.json")        # dram bytes per launch from the committed ncu capture
    if os.path.exists(tpath):
        with open(tpath) as f:
            for k, v in json.load(f).items():
                if k.startswith("fps_cluster_kernel"):
                    traffic = v["dram_bytes"]
    roofline = {"kernel": "fps_cluster_kernel (SA1: 16 clouds x 8192 pts -> 512 samples, 4-CTA clusters)", "bound": "hbm",
                "achieved": round(achieved, 1), "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": traffic,
                "algorithmic_bytes": fps_bytes, "kernel_us": round(fps_t * 1e6, 2),
                "note": "effective bytes B*(m-1)*N*16 (SURVEY 8d); data is register/SMEM resident, compulsory HBM is 1.6 MB"}

    cb = (cpu_baseline(steps=2, warmup=1, batch=2) if not os.environ.get('CPFN_BENCH_NO_CPU')
          else {'value': None, 'unit': UNIT, 'cores': 0, 'kind': 'port', 'sample': 'skipped'})
    total_points = world * B_PER_GPU * N_POINTS
    ms_per_step = dev_ms / args.steps
    line = {
        "metric": METRIC, "value": total_points / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16x3-split operands, f32 accumulate (MLPs); f32 (index ops); f32 sums / f64 solves (fitters)", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": B_PER_GPU, "n_points": N_POINTS, "k_slots": K_SLOTS,
                   "heads": [3, 4, K_SLOTS], "l2": "flushed between timed iterations (512 MB memset outside the event pairs)",
                   "sharding": "clouds sharded across ranks, no data-path collective", "fused_mlp": bool(fused.available()),
                   "cuda_graph": bool(use_graph)},
        "e2e": {"value": total_points / (e2e_ms * 1e-3 / args.steps), "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps,
                "api": "GlobalSPFN.stream_host (two batches in flight; every step does its full H2D, forward, fit, D2H)"
                       if pipelined else "GlobalSPFN.run_host",
                "single_call_latency_ms": lat_ms / args.steps},
        "gpu_launches": launches, "wall_ms_per_step_incl_flush": t_wall * 1e3 / args.steps,
        "fits_per_s": 4 * world * B_PER_GPU * K_SLOTS / (ms_per_step * 1e-3),
        "breakdown_us": breakdown, "roofline": roofline,
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "clocks": clocks,
    }
    _emit(line)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


_RESULT_FD = None


def _claim_stdout():
    """Libraries (NCCL's version banner, for one) write to file descriptor 1.  The contract is ONE JSON line on
    stdout: keep a private duplicate of the real stdout for that line and point fd 1 at stderr for everything else."""
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)


def _emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        args.steps = min(args.steps, 20)
        args.warmup = min(args.warmup, 2)
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
