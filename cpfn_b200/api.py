"""GlobalSPFN / LocalSPFN inference engine: the public entry point of the hot path.

``GlobalSPFN.forward(P)`` = PointNet2 forward (FPS + ball query + SA / FP layers + heads,
reference PointNet2/pn2_network.py:38-73) -> X normalised, W soft-maxed
(Utils/training_utils.py:141-142) -> ``compute_parameters`` for the four primitive types
(SPFN/losses_implementation.py:255-278).  ``run_host`` is the same call on HOST buffers
(pinned staging, host<->device copies included) -- what bench.py reports as ``e2e``.
"""
import os

import torch

from . import cuda_ops
from .pn2_network import PointNet2
from .spfn import losses_implementation as L


class GlobalSPFN:
    def __init__(self, output_sizes=(3, 4, 28), device="cuda:0", classes=('plane', 'sphere', 'cylinder', 'cone')):
        self.device = torch.device(device)
        self.classes = list(classes)
        self.output_sizes = list(output_sizes)
        self.model = PointNet2(dim_input=3, dim_pos=3, output_sizes=list(output_sizes)).to(self.device).eval()
        for p in self.model.parameters():
            p.requires_grad_(False)
        self._pin = {}
        self._graphs = {}
        self._sources = None
        self._graph_sig = None

    def load_state_dict(self, sd, strict=True):
        """Reference-layout state dict (training_SPFN.py:72-74 loads with strict=True)."""
        out = self.model.load_state_dict(sd, strict=strict)
        self.invalidate()
        return out

    def invalidate(self):
        """Forget the captured graphs and the packed weights.  Weight changes are detected without this (every
        replay compares the parameters' version counters with the ones the graphs were captured from); call it
        after REPLACING parameter objects of ``self.model``."""
        from . import fused
        fused.invalidate(self.model)
        self._graphs.clear()
        self._sources = None

    def _weights_changed(self):
        """True when a parameter / BatchNorm statistic of the model was written since the graphs were captured
        (their packed weight images are baked into the graphs)."""
        from . import fused
        if self._sources is None:
            self._sources = fused.model_sources(self.model)
        sig = fused.signature(self._sources)
        changed = self._graph_sig is not None and sig != self._graph_sig
        self._graph_sig = sig
        return changed

    @torch.no_grad()
    def forward(self, P, dropout=True, fit=True, scatter=None, rng_slot=0):
        """P [B,N,3] float32 on the device.  Returns a dict: X [B,N,3] unit normals, T [B,N,n_types]
        type logits, W [B,N,K] soft memberships, X_raw/T_raw/W_raw head outputs, l3_feats,
        output_feat, sa1_fps (int32 [B,512]) and ``parameters`` (the reference's dictionary).
        ``dropout=False`` replaces the reference's always-on dropout by the identity (parity runs).
        ``scatter``: see fused.spfn_post -- X / W / T are written into slabs of the given (possibly peer-mapped)
        arrays instead of fresh tensors (patch-sharded cascade)."""
        from . import fused
        if fused.available():
            if not P.is_cuda:
                raise RuntimeError("CPU not supported")
            heads, l3_feats, feat, l1_xyz, l2_xyz, packed = fused.pointnet2_forward(self.model, P, dropout=dropout,
                                                                                    rng_slot=rng_slot)
            out = {"heads": list(heads), "l3_feats": l3_feats, "output_feat": feat,
                   "l1_pos": l1_xyz.permute(0, 2, 1), "l2_pos": l2_xyz.permute(0, 2, 1)}
            if len(heads) != 3:
                # not the SPFN head layout (PatchSelection: one head of two logits, evaluation_PatchSelection.py:45):
                # the raw heads are the result, there is nothing to fit
                out["X_raw"] = heads[0]
                return out
            out.update({"X_raw": heads[0], "T_raw": heads[1], "W_raw": heads[2]})
            if heads[0].shape[2] == 3 and heads[2].shape[2] <= 64:
                # SPFN post-processing (Utils/training_utils.py:141-142) in one kernel
                nt = heads[1].shape[2]
                out["X"], out["W"], out["instance"], out["type"] = fused.spfn_post(packed, 0, 3 + nt, heads[2].shape[2],
                                                                                   t_off=3, n_types=nt, scatter=scatter)
            elif scatter is not None:
                raise RuntimeError("scatter needs the SPFN head layout [3, n_types, K <= 64]")
            else:
                out["X"] = torch.nn.functional.normalize(heads[0], p=2, dim=2, eps=1e-12)
                out["W"] = torch.softmax(heads[2], dim=2)
            out["T"] = heads[1]
            if fit:
                out["parameters"], out["parameters_packed"] = L.compute_parameters_packed(P, out["W"], out["X"], self.classes)
            return out
        raise RuntimeError("cpfn_b200: the fused CUDA path is unavailable (no CUDA device, or libcpfn_b200.so lacks "
                           "cpfn_mlp_chain); there is no fallback")

    def _capture(self, fn):
        """Warm ``fn`` up on a side stream (packs weights, sizes workspaces, sets kernel attributes),
        then capture it.  Returns (graph, fn's result, launches of this library inside the graph)."""
        cur = torch.cuda.current_stream(self.device)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(2):
                fn()
        cur.wait_stream(side)
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        n0 = cuda_ops.LAUNCHES
        with torch.cuda.graph(graph):
            out = fn()
        return graph, out, cuda_ops.LAUNCHES - n0

    @torch.no_grad()
    def forward_graphed(self, P, dropout=True, fit=True, scatter=None):
        """``forward`` replayed from a CUDA graph (captured once per input shape): the ~50 kernel
        launches of a step become one graph launch.  ``P`` may be a device tensor or a pinned host
        tensor (copied straight into the graph's input).  The returned tensors are STATIC buffers that
        the next call overwrites; the always-on dropout still draws a fresh mask per call (torch's
        graph-safe generator)."""
        graph, static_in, out, n_launch = self._net_graph(P, dropout, fit, 0, scatter=scatter)
        static_in.copy_(P, non_blocking=True)
        if dropout:
            self._sync_rng(P.shape)
        graph.replay()
        cuda_ops.count_launches(n_launch)
        return out

    def _sync_rng(self, shape, slot=0):
        """A captured forward draws its dropout mask from the device-side RNG state (one per graph slot): refresh
        it from torch's generator before every replay (fused.sync_rng)."""
        from . import fused
        fused.sync_rng(self.device, shape[0] * 128 * shape[1], slot)

    def _net_graph(self, P, dropout, fit, slot, scatter=None):
        """(graph, its static input, its static outputs, launches) for inputs shaped like P; ``slot`` selects one
        of several independent copies (own input / output buffers) so that consecutive batches can overlap."""
        if self._weights_changed():
            self._graphs.clear()
        key = (tuple(P.shape), bool(dropout), bool(fit)) + ((slot,) if slot else ())
        if scatter is not None:                               # the destination pointers are baked into the graph
            key += (("scatter", scatter["W"].data_ptr(), scatter["X"].data_ptr(), scatter["T"].data_ptr(),
                     int(scatter["stride"]), int(scatter["offset"])),)
        entry = self._graphs.get(key)
        if entry is None:
            static_in = torch.empty(tuple(P.shape), dtype=torch.float32, device=self.device)
            static_in.copy_(P)
            graph, out, n = self._capture(lambda: self.forward(static_in, dropout=dropout, fit=fit, scatter=scatter,
                                                               rng_slot=slot))
            entry = (graph, static_in, out, n)
            self._graphs[key] = entry
        return entry

    @torch.no_grad()
    def _fit_graphed(self, static_in, net_out):
        """The four fitters on the network graph's static outputs, as a second graph: run_host starts the
        device->host copies of the per-point results between the two, so they overlap the fitters."""
        key = ("fit", static_in.data_ptr(), net_out["W"].data_ptr())
        entry = self._graphs.get(key)
        if entry is None:
            graph, res, n = self._capture(
                lambda: L.compute_parameters_packed(static_in, net_out["W"], net_out["X"], self.classes))
            entry = (graph, res, n)
            self._graphs[key] = entry
        graph, res, n_launch = entry
        graph.replay()
        cuda_ops.count_launches(n_launch)
        return res

    @torch.no_grad()
    def stream_host(self, batches, dropout=True):
        """``run_host`` over a sequence of pinned host batches of ONE shape with several batches in flight (see
        ``_stream``): while a batch is on the SMs, the next ones' clouds are already crossing PCIe into their own
        graph buffers, and its per-point results travel back under its own fitters.  Every batch still does its full
        H2D, forward, fit and D2H; only their overlap changes.  Yields (results, h2d_bytes, d2h_bytes) in order; the
        result tensors of a batch are pinned buffers that are reused CPFN_LANES batches later."""
        yield from self._stream(batches, dropout, host=True)

    @torch.no_grad()
    def stream_device(self, batches, dropout=True):
        """The same pipeline for batches that already live on the device: yields (forward dictionary incl.
        ``parameters``, 0, 0) per batch.  The tensors are the static buffers of the batch's graph slot: valid until
        CPFN_LANES further batches have been submitted."""
        yield from self._stream(batches, dropout, host=False)

    def _stream(self, batches, dropout, host):
        """Several graph slots ("lanes", 6 by default, CPFN_LANES), each with its own stream, input / output buffers
        and RNG state.  Consecutive batches rotate through the lanes, so a batch's sampling -- a chain of dependent
        rounds that keeps 64 of the 148 SMs busy -- and the latency-bound hand-offs inside its MLP chains run beside
        the other batches' work instead of in front of it (the GPU interleaves the graphs; each graph is unchanged
        and so are its results)."""
        from . import fused
        dev = self.device
        caller = torch.cuda.current_stream(dev)
        n_lanes = max(1, int(os.environ.get("CPFN_LANES", "6")))       # measured: 2 -> 0.734, 4 -> 0.632, 6 -> 0.611, 8 -> 0.617 ms
        if len(self.__dict__.get("_lanes", ())) != n_lanes:
            self._lanes = [torch.cuda.Stream(device=dev) for _ in range(n_lanes)]
        lanes = self._lanes
        copy_in = self.__dict__.setdefault("_copy_in", torch.cuda.Stream(device=dev))
        copier = fused._side_stream(dev)
        free = [None] * n_lanes                               # slot's input may be overwritten after this event
        sent = [None] * n_lanes                               # slot's per-point outputs have left the device
        inflight = []
        start = torch.cuda.Event()
        start.record(caller)
        for lane in lanes:
            lane.wait_event(start)                            # whatever the caller enqueued before comes first
        for i, P_in in enumerate(batches):
            if host and not P_in.is_pinned():
                raise RuntimeError("stream_host needs pinned host tensors (torch.Tensor.pin_memory())")
            if not host and not P_in.is_cuda:
                raise RuntimeError("stream_device needs CUDA tensors")
            slot = i % n_lanes
            lane = lanes[slot]
            graph, static_in, out, n_launch = self._net_graph(P_in, dropout, False, slot)
            if "instance" not in out or self.classes != ['plane', 'sphere', 'cylinder', 'cone']:
                raise RuntimeError("streaming supports the SPFN head layout [3, n_types, K <= 64] with all four fitters")
            if host:
                with torch.cuda.stream(copy_in):              # H2D on its own stream, as early as the slot allows
                    if free[slot] is not None:
                        copy_in.wait_event(free[slot])
                    static_in.copy_(P_in, non_blocking=True)
                    arrived = torch.cuda.Event()
                    arrived.record(copy_in)
                lane.wait_event(arrived)
            res, tag = {}, "s%d_" % slot
            with torch.cuda.stream(lane):
                if not host:
                    static_in.copy_(P_in, non_blocking=True)
                if sent[slot] is not None:
                    lane.wait_event(sent[slot])               # the graph overwrites what that copy reads
                if dropout:
                    self._sync_rng(P_in.shape, slot)
                graph.replay()
                cuda_ops.count_launches(n_launch)
                ready = torch.cuda.Event()
                ready.record(lane)
                copied = ready
                if host:
                    with torch.cuda.stream(copier):           # per-point results go home while the fitters run
                        copier.wait_event(ready)
                        for k, v in (("normals", out["X"]), ("instance", out["instance"]), ("type", out["type"])):
                            res[k] = self._pinned(tag + k, v.shape, v.dtype)
                            res[k].copy_(v, non_blocking=True)
                        copied = torch.cuda.Event()
                        copied.record(copier)
                params, packed = self._fit_graphed(static_in, out)
                if host:
                    hp = self._pinned(tag + "params", packed.shape, packed.dtype)
                    hp.copy_(packed, non_blocking=True)
                    o = 0
                    for k, v in params.items():
                        res[k] = hp[o:o + v.numel()].view(v.shape)
                        o += v.numel()
                else:
                    res = dict(out)
                    res["parameters"], res["parameters_packed"] = params, packed
                done = torch.cuda.Event()
                done.record(lane)
            free[slot], sent[slot] = done, copied             # the fitters were the last readers of static_in
            h2d = P_in.numel() * 4 if host else 0
            d2h = (packed.numel() * 4 + sum(res[k].numel() * res[k].element_size() for k in ("normals", "instance", "type"))
                   if host else 0)
            inflight.append((res, done, copied, h2d, d2h))
            if len(inflight) >= n_lanes:                      # hand out the oldest batch while the newer ones run
                r = inflight.pop(0)
                r[1].synchronize(); r[2].synchronize()
                yield r[0], r[3], r[4]
        for r in inflight:
            r[1].synchronize(); r[2].synchronize()
            yield r[0], r[3], r[4]
        for lane in lanes:
            caller.wait_stream(lane)

    def _pinned(self, name, shape, dtype):
        t = self._pin.get(name)
        if t is None or t.shape != tuple(shape) or t.dtype != dtype:
            t = torch.empty(shape, dtype=dtype, pin_memory=True)
            self._pin[name] = t
        return t

    @torch.no_grad()
    def run_host(self, P_host, dropout=True, graphed=False, overlap_d2h=True):
        """P_host: CPU float32 [B,N,3] (pinned or pageable).  Copies it to the device, runs
        forward + fitters, and returns HOST tensors: the parameter dictionary, per-point
        instance labels (int32 [B,N], argmax of W), per-point type labels and unit normals --
        what evaluation_globalSPFN.py:97-110 of the reference saves per shape.
        Returns (results, h2d_bytes, d2h_bytes)."""
        B, N, _ = P_host.shape
        if graphed and overlap_d2h and self.classes == ['plane', 'sphere', 'cylinder', 'cone']:
            src = P_host if P_host.is_pinned() else self._pinned("P", P_host.shape, torch.float32).copy_(P_host)
            out = self.forward_graphed(src, dropout=dropout, fit=False)
            if "instance" in out:
                main = torch.cuda.current_stream(self.device)
                from . import fused
                copier = fused._side_stream(self.device)
                ready = torch.cuda.Event()
                ready.record(main)
                res = {}
                with torch.cuda.stream(copier):          # per-point results go home while the fitters run
                    copier.wait_event(ready)
                    for k, v in (("normals", out["X"]), ("instance", out["instance"]), ("type", out["type"])):
                        res[k] = self._pinned("out_" + k, v.shape, v.dtype)
                        res[k].copy_(v, non_blocking=True)
                params, packed = self._fit_graphed(self._graphs[(tuple(P_host.shape), bool(dropout), False)][1], out)
                hp = self._pinned("out_params", packed.shape, packed.dtype)
                hp.copy_(packed, non_blocking=True)
                o = 0
                for k, v in params.items():
                    res[k] = hp[o:o + v.numel()].view(v.shape)
                    o += v.numel()
                d2h = packed.numel() * 4 + sum(res[k].numel() * res[k].element_size() for k in ("normals", "instance", "type"))
                main.synchronize()
                copier.synchronize()
                return res, P_host.numel() * 4, d2h
        if P_host.is_pinned():
            stage = P_host
        else:
            stage = self._pinned("P", P_host.shape, torch.float32)
            stage.copy_(P_host)
        if graphed:
            out = self.forward_graphed(stage, dropout=dropout)           # H2D straight into the graph's input
        else:
            out = self.forward(stage.to(self.device, non_blocking=True), dropout=dropout)
        params = out["parameters"]
        res_dev = {"instance": out["instance"] if "instance" in out else torch.argmax(out["W"], dim=2).to(torch.int32),
                   "type": out["type"] if "type" in out else torch.argmax(out["T"], dim=2).to(torch.int32),
                   "normals": out["X"]}
        packed = out.get("parameters_packed")
        if packed is not None and packed.numel() != sum(v.numel() for v in params.values()):
            packed = None                           # a subset of the classes was requested: copy tensor by tensor
        res, d2h = {}, 0
        if packed is not None:                     # the ten parameter tensors are views of ONE buffer: one copy
            hp = self._pinned("out_params", packed.shape, packed.dtype)
            hp.copy_(packed, non_blocking=True)
            d2h += packed.numel() * 4
            o = 0
            for k, v in params.items():
                res[k] = hp[o:o + v.numel()].view(v.shape)
                o += v.numel()
        else:
            res_dev.update(params)
        for k, v in res_dev.items():
            h = self._pinned("out_" + k, v.shape, v.dtype)
            h.copy_(v, non_blocking=True)
            res[k] = h
            d2h += v.numel() * v.element_size()
        torch.cuda.current_stream(self.device).synchronize()
        return res, P_host.numel() * 4, d2h


class LocalSPFN:
    """One shape through the LocalSPFN inference path of evaluation_localSPFN.py: patches of the high-resolution
    cloud (the k nearest points of every seed, Utils/sampling_utils.py) -> per-patch normalisation
    (Dataset/dataloaders.py:249-253) -> the PointNet2 backbone with K = n_max_local_instances slots on the patches
    as the batch dimension (:95-97) -> patch-to-object merging with the object-level prediction
    (:99-130, Utils/merging_utils.py).  Everything between the seeds and the merged result stays on the device;
    the greedy label merge is the one host step (as in the reference)."""

    def __init__(self, n_max_local_instances=21, n_types=4, device="cuda:0", num_points_patch=8192):
        self.engine = GlobalSPFN(output_sizes=(3, n_types, n_max_local_instances), device=device)
        self.device = self.engine.device
        self.num_points_patch = int(num_points_patch)

    def load_state_dict(self, sd, strict=True):
        return self.engine.load_state_dict(sd, strict=strict)

    def _peer_exchange(self, nb, Np, group, merge_rank):
        """Receive buffers of the patch-sharded cascade in SYMMETRIC memory (torch.distributed._symmetric_memory: one
        allocation per rank, every rank's copy mapped into every other rank's address space over NVLink): W [nb,Np,Kl]
        | X [nb,Np,3] | T [nb,Np,n_types] float32 and the int64 patch indices [nb,Np].  Returns a dict with the
        merge rank's buffers as seen from THIS rank (peer-mapped unless this is the merge rank) and the handles for
        the device-side barriers; None when symmetric memory is unavailable (the NCCL all-gather path is used)."""
        import torch.distributed as dist
        key = (nb, Np, merge_rank, id(group))
        cache = self.__dict__.setdefault("_exchanges", {})
        if key in cache:
            return cache[key]
        ex = None
        try:
            import torch.distributed._symmetric_memory as symm
            Kl, nt = self.engine.output_sizes[2], self.engine.output_sizes[1]
            g = group if group is not None else dist.group.WORLD
            n_f = nb * Np * (Kl + 3 + nt)
            fbuf = symm.empty(n_f, dtype=torch.float32, device=self.device)
            ibuf = symm.empty(nb * Np, dtype=torch.int64, device=self.device)
            fh, ih = symm.rendezvous(fbuf, g), symm.rendezvous(ibuf, g)
            views = {}
            for name, hdl_rank in (("dst", merge_rank), ("own", dist.get_rank(group))):
                o = 0
                v = {}
                for f, width in (("W", Kl), ("X", 3), ("T", nt)):
                    v[f] = fh.get_buffer(hdl_rank, (nb, Np, width), torch.float32, o)
                    o += nb * Np * width
                v["idx"] = ih.get_buffer(hdl_rank, (nb, Np), torch.int64, 0)
                views[name] = v
            ex = {"dst": views["dst"], "own": views["own"], "fh": fh, "ih": ih, "keep": (fbuf, ibuf)}
        except Exception as e:                                   # no peer mapping on this system: NCCL path
            self.__dict__["_exchange_error"] = "%s: %s" % (type(e).__name__, e)
        cache[key] = ex
        return ex

    @staticmethod
    def normalise_patches(P_global, patch_indices):
        """dataloaders.py:249-253: centre every patch on its mean and scale it into the unit ball -- one kernel
        (gather + mean + max norm + scale, csrc/glue.cu) whose per-patch result does not depend on how many patches
        are normalised together (a sharded run is bit-identical to a single-GPU one)."""
        from . import _lib
        if not P_global.is_cuda:
            raise RuntimeError("CPU not supported")
        P_global = P_global.to(torch.float32).contiguous()
        idx = patch_indices.contiguous()
        if idx.dtype not in (torch.int32, torch.int64):
            idx = idx.to(torch.int64)
        nb, Np = idx.shape
        out = torch.empty(nb, Np, 3, dtype=torch.float32, device=P_global.device)
        with torch.cuda.device(P_global.device):
            _lib.check(_lib.lib().cpfn_normalise_patches(P_global.data_ptr(), P_global.shape[0], idx.data_ptr(),
                                                         int(idx.dtype == torch.int64), nb, Np, out.data_ptr(),
                                                         torch.cuda.current_stream(P_global.device).cuda_stream),
                       "normalise_patches")
        cuda_ops.count_launches(1)
        return out

    @torch.no_grad()
    def run_shape(self, P_global, spfn_labels, spfn_normals, spfn_type, seeds=None, patch_indices=None, dropout=True,
                  graphed=False, threshold=0):
        """P_global [Ng,3]; object-level prediction spfn_labels [Ng,Kg], spfn_normals [Ng,3], spfn_type [Ng,n_types];
        either ``seeds`` [S,3] (patches are extracted here) or ``patch_indices`` int [nb,Np].  Returns a dict with
        W_fusion [Ng,L], X_global [Ng,3], T_global [Ng,n_types], labels (numpy int64) and patch_indices."""
        from . import merging_utils, sampling_utils
        P_global = P_global.to(self.device, torch.float32).contiguous()
        if patch_indices is None:
            if seeds is None:
                raise ValueError("run_shape needs seeds or patch_indices")
            patch_indices = sampling_utils.extract_patches(P_global, seeds.to(self.device, torch.float32),
                                                           self.num_points_patch)
        patch_indices = patch_indices.to(self.device)
        if patch_indices.shape[0] == 0:                       # evaluation_localSPFN.py:131-135: nothing to merge
            raise ValueError("run_shape needs at least one patch (the reference falls back to the object-level labels)")
        P = self.normalise_patches(P_global, patch_indices)
        out = (self.engine.forward_graphed if graphed else self.engine.forward)(P, dropout=dropout, fit=False)
        W_fusion, X_global, T_global, labels = merging_utils.merge_shape(
            out["W"], out["X"], out["T"], patch_indices, spfn_labels.to(self.device), spfn_normals.to(self.device),
            spfn_type.to(self.device), threshold=threshold)
        return {"W_fusion": W_fusion, "X_global": X_global, "T_global": T_global, "labels": labels,
                "patch_indices": patch_indices, "W": out["W"], "X": out["X"], "T": out["T"]}

    @torch.no_grad()
    def run_shape_sharded(self, P_global, spfn_labels, spfn_normals, spfn_type, seeds=None, patch_indices=None,
                          dropout=True, graphed=True, threshold=0, group=None, merge_rank=0, timings=None,
                          exchange="auto"):
        """``run_shape`` with the shape's patches sharded over the ranks of ``group``: every rank passes
        the SAME shape (high-resolution cloud, object-level prediction, seeds or patch indices); rank r extracts,
        normalises and runs the LocalSPFN backbone on patches r, r+G, ... only (replicated weights, no collective),
        the per-point outputs {W, X, T, patch indices} of all patches travel to rank ``merge_rank``, which runs the
        merge (evaluation_localSPFN.py:99-130; Utils/merging_utils.similarity_soft needs every patch's memberships).
        ``exchange``:
          "p2p"   the kernel that produces W / X / T (softmax / normalise, cpfn_spfn_post_scatter) writes them straight
                  into the merge rank's receive buffers -- symmetric memory, peer-mapped over NVLink -- so compute and
                  transfer are one kernel; two device-side barriers order it against the merge (no NCCL on the path);
          "nccl"  one all-gather per dtype (dist.all_gather_patches), then a row permutation;
          "auto"  "p2p" when symmetric memory is available and the patch size is a multiple of 256 points.
        Returns the ``run_shape`` dictionary on ``merge_rank`` (with "p2p" its W / X / T / patch_indices are views of
        the receive buffers, valid until the next call) and None elsewhere.  ``timings``: optional dict that receives
        CUDA events around the stages (bench.py)."""
        import torch.distributed as dist
        from . import dist as cdist, merging_utils, sampling_utils
        alone = not (dist.is_available() and dist.is_initialized())          # no process group: a world of one rank
        world, rank = (1, 0) if alone else (dist.get_world_size(group), dist.get_rank(group))
        dev = self.device
        P_global = P_global.to(dev, torch.float32).contiguous()
        nb = int(seeds.shape[0] if patch_indices is None else patch_indices.shape[0])
        if nb == 0:
            raise ValueError("run_shape_sharded needs at least one patch")
        Np = min(self.num_points_patch, P_global.shape[0]) if patch_indices is None else int(patch_indices.shape[1])
        mine = cdist.shard_units(nb, rank, world)
        sel = torch.tensor(mine, dtype=torch.int64, device=dev)
        Kl, nt = self.engine.output_sizes[2], self.engine.output_sizes[1]
        ex = None
        if world > 1 and exchange in ("auto", "p2p") and Np % 256 == 0:
            ex = self._peer_exchange(nb, Np, group, merge_rank)
        if exchange == "p2p" and world > 1 and ex is None:
            raise RuntimeError("peer exchange unavailable: %s" % self.__dict__.get("_exchange_error", "patch size"))

        def mark(name):
            if timings is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                timings[name] = e
        mark("start")
        if ex is not None:
            ex["fh"].barrier(channel=0)          # the previous shape's merge has consumed the receive buffers
        if len(mine):
            if patch_indices is None:
                my_idx = sampling_utils.extract_patches(P_global, seeds.to(dev, torch.float32).index_select(0, sel),
                                                        self.num_points_patch)
            else:
                my_idx = patch_indices.to(dev).index_select(0, sel)
            mark("extracted")
            P = self.normalise_patches(P_global, my_idx)
            run = self.engine.forward_graphed if graphed else self.engine.forward
            if ex is not None:
                ex["dst"]["idx"][rank::world][:len(mine)].copy_(my_idx)              # patch j -> slab j * G + rank
                run(P, dropout=dropout, fit=False,
                    scatter={"X": ex["dst"]["X"], "W": ex["dst"]["W"], "T": ex["dst"]["T"], "stride": world, "offset": rank})
            else:
                out = run(P, dropout=dropout, fit=False)
                if world > 1:
                    feats = torch.cat([out["W"], out["X"], out["T"]], dim=2)         # [b, Np, Kl + 3 + n_types]
        else:                                                                        # more ranks than patches
            my_idx = torch.empty(0, Np, dtype=torch.int64, device=dev)
            mark("extracted")
            feats = torch.empty(0, Np, Kl + 3 + nt, dtype=torch.float32, device=dev)
        mark("backbone")
        if ex is not None:
            ex["fh"].barrier(channel=1)          # every rank's writes have landed in the merge rank's memory
            own = ex["own"]
            W, X, T, idx_all = own["W"], own["X"], own["T"], own["idx"]
        elif world == 1:
            W, X, T, idx_all = out["W"], out["X"], out["T"], my_idx.to(torch.int64)
        else:
            feats_all, idx_all = cdist.all_gather_patches(feats, my_idx, nb, group=group)
            W, X, T = (feats_all[:, :, :Kl].contiguous(), feats_all[:, :, Kl:Kl + 3].contiguous(),
                       feats_all[:, :, Kl + 3:].contiguous())
        mark("gathered")
        if rank != merge_rank:
            return None
        W_fusion, X_global, T_global, labels = merging_utils.merge_shape(
            W, X, T, idx_all, spfn_labels.to(dev), spfn_normals.to(dev), spfn_type.to(dev), threshold=threshold)
        mark("merged")
        return {"W_fusion": W_fusion, "X_global": X_global, "T_global": T_global, "labels": labels,
                "patch_indices": idx_all, "W": W, "X": X, "T": T,
                "exchange": "p2p" if ex is not None else ("none" if world == 1 else "nccl")}
