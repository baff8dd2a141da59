"""numpy restatement of the segmentation glue of the SPFN losses (TEST INFRASTRUCTURE, NOT PRODUCT CODE).

  hungarian_matching   SPFN/losses_implementation.py:11-30 (scipy.optimize.linear_sum_assignment on the IoU matrix)
  compute_miou_loss    SPFN/losses_implementation.py:77-89

Parity pin: tests/golden/ref_seg.npz, outputs of the UNMODIFIED reference functions on CPU tensors
(tests/golden/make_ref_seg_golden.py).
"""
import numpy as np
from scipy.optimize import linear_sum_assignment

F = np.float32


def iou_cost(W_pred_b, I_gt_b):
    """:20-25 for one sample: (cost [K', K] float32, K')."""
    n_gt = int(I_gt_b.max()) + 1
    W_gt = np.eye(n_gt + 1, dtype=F)[I_gt_b]                      # -1 indexes the extra last column, as in the reference
    dot = (W_gt.T @ W_pred_b.astype(F)).astype(F)
    den = W_gt.sum(0, dtype=F)[:, None] + W_pred_b.astype(F).sum(0, dtype=F)[None, :] - dot
    return (dot / np.maximum(den, F(1e-10)))[:n_gt].astype(F), n_gt


def hungarian_matching(W_pred, I_gt, with_mask=False):
    B, N, K = W_pred.shape
    out = np.zeros((B, K), dtype=np.int64)
    mask = np.zeros((B, K), dtype=bool)
    for b in range(B):
        cost, n_gt = iou_cost(W_pred[b], I_gt[b])
        _, col = linear_sum_assignment(-cost)
        out[b, :n_gt] = col
        mask[b, :n_gt] = True
    return (out, mask) if with_mask else out


def compute_miou_loss(W, I_gt, matching_indices, div_eps=1e-10):
    B, N, K = W.shape
    n_labels = matching_indices.shape[1]
    W_re = np.take_along_axis(W.astype(F), np.broadcast_to(matching_indices[:, None, :], (B, N, n_labels)), axis=2)
    W_gt = np.eye(n_labels + 2, dtype=F)[I_gt][:, :, :n_labels]   # -1 -> last row -> sliced off: a zero row
    dot = np.sum(W_gt * W_re, axis=1, dtype=F)
    den = W_gt.sum(1, dtype=F) + W_re.sum(1, dtype=F) - dot
    return 1.0 - dot / (den + F(div_eps)), 1 - dot / N
