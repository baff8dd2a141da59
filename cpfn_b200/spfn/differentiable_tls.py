"""``solve_weighted_tls`` and the custom-SVD column of the reference
(SPFN/differentiable_tls.py:123-143, 200-209) on the CUDA moment kernels."""
import torch

from . import _reference, _train

guard_one_over_matrix = _train.guard_one_over_matrix


def compute_svd_K(s):
    """res[b,i,j] = 1/(s_i^2 - s_j^2) guarded, 0 on the diagonal (reference :45-53)."""
    s = s ** 2
    return guard_one_over_matrix(s.unsqueeze(2) - s.unsqueeze(1))


class Custom_svd_v_colum:
    """``Custom_svd_v_colum().apply(M)`` -> last right-singular vector of the symmetric M [B,n,n]."""

    @staticmethod
    def apply(M, col_index=-1):
        if col_index != -1:
            raise NotImplementedError("only the last column is used by SPFN")
        return _train.svd_v_last_column(M.double()).to(M.dtype)


def solve_weighted_tls(A, W):
    """A [B',N,3], W [B',N] -> x [B',3]: argmin_{|x|=1} x^T (sum_n w a a^T) x.  Differentiable w.r.t. W and A."""
    M = _train.weighted_moments(W.unsqueeze(2), A.detach(), A)          # x x^T features carry A's gradient
    Mxx = _train._sym3(M[..., _train._X2:_train._X2 + 6]).squeeze(1)
    return _train.svd_v_last_column(Mxx).to(torch.float32)


__getattr__ = _reference.forwarder(globals(), "differentiable_tls", ('guard_one_over_matrix', 'compute_svd_K', 'Custom_svd_v_colum', 'solve_weighted_tls'))
