"""Data-parallel training step of the SPFN networks on this package's kernels (BASELINE configs[3]: LocalSPFN
forward / backward over the patches of a batch, sharded across the GPUs of one box).

The step is the reference's (Utils/training_utils.py:84-158): ``PointNet2`` in train mode -- the per-op CUDA kernels of
this package with their scatter-add backward (grouping, interpolation) around torch's convolution / batch-norm modules
-- normalise / soft-max, ``compute_all_losses`` (SPFN/losses_implementation.py:675-720), backward, optimiser step.
``compute_all_losses`` and the normal / type losses are loss glue outside the hot path: they are the reference's own
functions, reached through ``cpfn_b200.spfn.losses_implementation`` (the reference checkout must be on sys.path), with
the hot parts inside them -- Hungarian matching, the mIoU sums, the fitters, the residues -- replaced by this package's
kernels.  Sharding: rank r takes samples r, r+G, ... of the batch (cpfn_b200.dist.shard_units); weights are replicated
and the gradients are averaged with ONE NCCL all-reduce of a flat bucket (5.6 MB).  BatchNorm statistics stay per rank
(the reference is single-GPU)."""
import torch

from .spfn import losses_implementation as L

LOCAL_MULTIPLIERS = dict(normal_loss_multiplier=1.0, type_loss_multiplier=1.0, miou_loss_multiplier=1.0,
                         residue_loss_multiplier=0.0, parameter_loss_multiplier=0.0, total_loss_multiplier=1.0)
GLOBAL_MULTIPLIERS = dict(LOCAL_MULTIPLIERS, residue_loss_multiplier=1.0, parameter_loss_multiplier=1.0)


def forward_losses(model, batch, multipliers=LOCAL_MULTIPLIERS, classes=('plane', 'sphere', 'cylinder', 'cone')):
    """batch: dict with P [B,N,3], X_gt [B,N,3], I_gt int64 [B,N], T_gt int64 [B,K], points_per_instance
    [B,K,M,3], gt_parameters {plane_normal, cylinder_axis, cone_axis: [B,K,3]}.  Returns the tuple of
    compute_all_losses (total loss first)."""
    X, T, W, _, _ = model(batch["P"])
    X = torch.nn.functional.normalize(X, p=2, dim=2, eps=1e-12)
    W = torch.softmax(W, dim=2)
    m = multipliers
    return L.compute_all_losses(batch["P"], W, batch["I_gt"], X, batch["X_gt"], T, batch["T_gt"], batch["gt_parameters"],
                                batch["points_per_instance"], m["normal_loss_multiplier"], m["type_loss_multiplier"],
                                m["miou_loss_multiplier"], m["residue_loss_multiplier"], m["parameter_loss_multiplier"],
                                m["total_loss_multiplier"], False, mode_seg='mIoU', classes=list(classes))


def shard_batch(batch, rank, world):
    """This rank's samples of a batch held identically by every rank."""
    from . import dist as cdist
    B = batch["P"].shape[0]
    sel = torch.tensor(cdist.shard_units(B, rank, world), dtype=torch.int64, device=batch["P"].device)
    out = {k: (v.index_select(0, sel) if torch.is_tensor(v) else v) for k, v in batch.items() if k != "gt_parameters"}
    out["gt_parameters"] = {k: v.index_select(0, sel) for k, v in batch["gt_parameters"].items()}
    return out


def all_reduce_gradients(model, group=None, bucket=None):
    """Average the gradients over the ranks with one collective on a flat bucket."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    params = [p for p in model.parameters() if p.requires_grad and p.grad is not None]
    if bucket is None or bucket.numel() != sum(p.numel() for p in params):
        bucket = torch.empty(sum(p.numel() for p in params), dtype=torch.float32, device=params[0].device)
    o = 0
    for p in params:
        bucket[o:o + p.numel()].copy_(p.grad.reshape(-1))
        o += p.numel()
    dist.all_reduce(bucket, group=group)
    bucket.div_(world)
    o = 0
    for p in params:
        p.grad.copy_(bucket[o:o + p.numel()].view_as(p.grad))
        o += p.numel()
    return bucket


def train_step(model, optimizer, batch, multipliers=LOCAL_MULTIPLIERS, group=None, timings=None):
    """One optimisation step on this rank's share of ``batch`` (already sharded).  Returns the tuple of
    compute_all_losses.  ``timings``: optional dict receiving CUDA events (forward / backward / all-reduce / step)."""
    import torch.distributed as dist

    def mark(name):
        if timings is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            timings[name] = e
    model.train()
    optimizer.zero_grad()
    mark("start")
    losses = forward_losses(model, batch, multipliers)
    mark("forward")
    losses[0].backward()
    mark("backward")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        all_reduce_gradients(model, group)
    mark("all_reduce")
    optimizer.step()
    mark("step")
    return losses
