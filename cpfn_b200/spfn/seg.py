"""Segmentation glue of the SPFN losses / metrics on the device (csrc/seg_loss.cu):
``hungarian_matching`` (SPFN/losses_implementation.py:11-30, SPFN/metric_implementation.py:9-30) without the per-sample
device->host copy and scipy call, and ``compute_miou_loss`` (SPFN/losses_implementation.py:77-89) from the same sums,
differentiable w.r.t. W.  No CPU path: CPU tensors raise."""
import torch

from .. import _lib, cuda_ops


def _sums(W, I_gt, G):
    """(S [B,G,K], colsum [B,K], count [B,G], n_gt int32 [B]) -- see cpfn_label_membership_sums."""
    if not W.is_cuda:
        raise RuntimeError("CPU not supported")
    B, N, K = W.shape
    Wc = W.detach().float().contiguous()
    I = I_gt.contiguous()
    if I.dtype not in (torch.int32, torch.int64):
        I = I.to(torch.int64)
    dev = W.device
    S = torch.empty(B, G, K, dtype=torch.float32, device=dev)
    colsum = torch.empty(B, K, dtype=torch.float32, device=dev)
    count = torch.empty(B, G, dtype=torch.float32, device=dev)
    n_gt = torch.empty(B, dtype=torch.int32, device=dev)
    L = _lib.lib()
    with torch.cuda.device(dev):
        nbytes = L.cpfn_seg_workspace_bytes(B, N, K, G)
        if nbytes == 0:
            raise RuntimeError("segmentation sums support up to 64 instance slots")
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        _lib.check(L.cpfn_label_membership_sums(Wc.data_ptr(), I.data_ptr(), int(I.dtype == torch.int64), B, N, K, G,
                                                S.data_ptr(), colsum.data_ptr(), count.data_ptr(), n_gt.data_ptr(),
                                                ws.data_ptr(), nbytes, torch.cuda.current_stream(dev).cuda_stream),
                   "label_membership_sums")
    cuda_ops.count_launches(2)
    return S, colsum, count, n_gt


def hungarian_matching(W_pred, I_gt, with_mask=False):
    """W_pred [B,N,K] float, I_gt [B,N] int (-1 = background) -> matching_indices int64 [B,K]: ground-truth primitive
    k of sample b is matched with predicted slot matching_indices[b,k]; entries beyond the sample's number of
    ground-truth labels are 0.  ``with_mask`` also returns the bool [B,K] mask of the meaningful entries
    (metric_implementation.hungarian_matching).  Nothing is copied to the host and nothing synchronises."""
    B, N, K = W_pred.shape
    if K > 32:
        raise RuntimeError("device hungarian_matching supports up to 32 instance slots")
    S, colsum, count, n_gt = _sums(W_pred, I_gt, K)
    dev = W_pred.device
    matching = torch.empty(B, K, dtype=torch.int64, device=dev)
    mask = torch.empty(B, K, dtype=torch.uint8, device=dev) if with_mask else None
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().cpfn_hungarian_matching(S.data_ptr(), colsum.data_ptr(), count.data_ptr(), n_gt.data_ptr(),
                                                      B, K, K, matching.data_ptr(),
                                                      mask.data_ptr() if mask is not None else None,
                                                      torch.cuda.current_stream(dev).cuda_stream), "hungarian_matching")
    cuda_ops.count_launches(1)
    return (matching, mask.bool()) if with_mask else matching


class _MembershipSums(torch.autograd.Function):
    """(W [B,N,K], I_gt [B,N]) -> (S [B,G,K], colsum [B,K], count [B,G]); linear in W."""

    @staticmethod
    def forward(ctx, W, I_gt, G):
        S, colsum, count, _ = _sums(W, I_gt, G)
        ctx.save_for_backward(I_gt)
        ctx.G = G
        ctx.mark_non_differentiable(count)
        return S, colsum, count

    @staticmethod
    def backward(ctx, gS, gcol, _gcount):
        (I_gt,) = ctx.saved_tensors
        B, N = I_gt.shape
        G, K = ctx.G, gS.shape[2]
        # dW[b,n,k] = gS[b, I[b,n], k] (labelled points) + gcol[b,k]
        valid = (I_gt >= 0) & (I_gt < G)
        rows = torch.gather(gS, 1, I_gt.clamp(0, G - 1).to(torch.int64).unsqueeze(2).expand(B, N, K))
        return rows * valid.unsqueeze(2) + gcol.unsqueeze(1), None, None


def compute_miou_loss(W, I_gt, matching_indices, div_eps=1e-10):
    """SPFN/losses_implementation.py:77-89.  W [B,N,K], I_gt [B,N], matching_indices int64 [B,K'] ->
    (1 - mIoU [B,K'], 1 - dot / N [B,K']).  The O(B N K) sums are one kernel; the rest is [B,K]-sized algebra."""
    B, N, K = W.shape
    n_labels = matching_indices.shape[1]
    S, colsum, count = _MembershipSums.apply(W, I_gt, n_labels)
    m = matching_indices.to(torch.int64)
    dot = torch.gather(S, 2, m.unsqueeze(2)).squeeze(2)                 # sum_n [I = k] W[n, m_k]
    denominator = count + torch.gather(colsum, 1, m) - dot
    mIoU = dot / (denominator + div_eps)
    return 1.0 - mIoU, 1 - dot / N
