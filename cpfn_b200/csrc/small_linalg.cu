// Batched tiny dense linear algebra in fp64 for the differentiable (training) path of the SPFN fitters
// (SPFN/differentiable_tls.py:123-143 Custom_svd_v_colum, SPFN/geometry_utils.py:121-142 guarded_matrix_solve_ls):
// the reference calls torch.svd / torch.solve on [B*K, 3, 3]-sized batches (cuSOLVER / MAGMA launches with host
// synchronisation); here one thread owns one matrix.
//   cpfn_sym_eigh_small   symmetric D x D (D = 2, 3): cyclic Jacobi, eigenvalues ascending, eigenvectors in the
//                         columns of Q (the contract of torch.linalg.eigh)
//   cpfn_small_solve      A x = b (or A^T x = b), D <= 3: Gaussian elimination with partial pivoting (LU, as
//                         torch.linalg.solve); a singular pivot yields inf / nan exactly like the library call
#include <math.h>

#include "common.cuh"

namespace cpfn {
namespace {

constexpr int kLinThreads = 128;

template <int D>
__global__ void __launch_bounds__(kLinThreads)
sym_eigh_kernel(const double *__restrict__ A, long long n, double *__restrict__ lam, double *__restrict__ Q) {
  const long long i = static_cast<long long>(blockIdx.x) * kLinThreads + threadIdx.x;
  if (i >= n) return;
  double a[D][D], v[D][D];
#pragma unroll
  for (int r = 0; r < D; ++r)
#pragma unroll
    for (int c = 0; c < D; ++c) {
      // symmetrise: the callers build M from symmetric moments, rounding may differ in the last bit
      a[r][c] = 0.5 * (A[i * D * D + r * D + c] + A[i * D * D + c * D + r]);
      v[r][c] = r == c ? 1.0 : 0.0;
    }
  for (int sweep = 0; sweep < 32; ++sweep) {
    double off = 0.0, diag = 0.0;
#pragma unroll
    for (int r = 0; r < D; ++r) {
      diag += a[r][r] * a[r][r];
#pragma unroll
      for (int c = r + 1; c < D; ++c) off += a[r][c] * a[r][c];
    }
    if (off <= 1e-300 || off <= 1e-34 * diag) break;          // off-diagonal below fp64 resolution of the spectrum
#pragma unroll
    for (int p = 0; p < D; ++p)
#pragma unroll
      for (int q = p + 1; q < D; ++q) {
        const double apq = a[p][q];
        if (apq == 0.0) continue;
        const double theta = (a[q][q] - a[p][p]) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        a[p][p] -= t * apq;
        a[q][q] += t * apq;
        a[p][q] = a[q][p] = 0.0;
#pragma unroll
        for (int r = 0; r < D; ++r) {
          if (r != p && r != q) {
            const double arp = a[r][p], arq = a[r][q];
            a[r][p] = a[p][r] = c * arp - s * arq;
            a[r][q] = a[q][r] = s * arp + c * arq;
          }
          const double vrp = v[r][p], vrq = v[r][q];
          v[r][p] = c * vrp - s * vrq;
          v[r][q] = s * vrp + c * vrq;
        }
      }
  }
  // ascending eigenvalues (selection sort of D <= 3 columns)
  int order[D];
#pragma unroll
  for (int r = 0; r < D; ++r) order[r] = r;
#pragma unroll
  for (int r = 0; r < D; ++r)
#pragma unroll
    for (int c = r + 1; c < D; ++c)
      if (a[order[c]][order[c]] < a[order[r]][order[r]]) { const int t = order[r]; order[r] = order[c]; order[c] = t; }
#pragma unroll
  for (int c = 0; c < D; ++c) {
    lam[i * D + c] = a[order[c]][order[c]];
    if (Q != nullptr) {
#pragma unroll
      for (int r = 0; r < D; ++r) Q[i * D * D + r * D + c] = v[r][order[c]];
    }
  }
}

template <int D>
__global__ void __launch_bounds__(kLinThreads)
small_solve_kernel(const double *__restrict__ A, const double *__restrict__ b, long long n, int transpose,
                   double *__restrict__ x) {
  const long long i = static_cast<long long>(blockIdx.x) * kLinThreads + threadIdx.x;
  if (i >= n) return;
  double m[D][D + 1];
#pragma unroll
  for (int r = 0; r < D; ++r) {
#pragma unroll
    for (int c = 0; c < D; ++c) m[r][c] = transpose ? A[i * D * D + c * D + r] : A[i * D * D + r * D + c];
    m[r][D] = b[i * D + r];
  }
#pragma unroll
  for (int k = 0; k < D; ++k) {
    int piv = k;
#pragma unroll
    for (int r = k + 1; r < D; ++r)
      if (fabs(m[r][k]) > fabs(m[piv][k])) piv = r;
#pragma unroll
    for (int r = k + 1; r < D; ++r)
      if (r == piv) {
#pragma unroll
        for (int c = 0; c <= D; ++c) { const double t = m[k][c]; m[k][c] = m[r][c]; m[r][c] = t; }
      }
    const double inv = 1.0 / m[k][k];
#pragma unroll
    for (int r = k + 1; r < D; ++r) {
      const double f = m[r][k] * inv;
#pragma unroll
      for (int c = k; c <= D; ++c) m[r][c] -= f * m[k][c];
    }
  }
  double sol[D];
#pragma unroll
  for (int r = D - 1; r >= 0; --r) {
    double s = m[r][D];
#pragma unroll
    for (int c = r + 1; c < D; ++c) s -= m[r][c] * sol[c];
    sol[r] = s / m[r][r];
  }
#pragma unroll
  for (int r = 0; r < D; ++r) x[i * D + r] = sol[r];
}

}  // namespace
}  // namespace cpfn

extern "C" int cpfn_sym_eigh_small(const double *A, long long n, int D, double *lam, double *Q, cpfn_stream_t stream) {
  using namespace cpfn;
  if (n < 0 || (D != 2 && D != 3)) return CPFN_EINVAL;
  if (n == 0) return CPFN_OK;
  if (!A || !lam) return CPFN_EINVAL;
  const unsigned grid = static_cast<unsigned>((n + kLinThreads - 1) / kLinThreads);
  if (D == 2) sym_eigh_kernel<2><<<grid, kLinThreads, 0, as_stream(stream)>>>(A, n, lam, Q);
  else sym_eigh_kernel<3><<<grid, kLinThreads, 0, as_stream(stream)>>>(A, n, lam, Q);
  return check_launch();
}

extern "C" int cpfn_small_solve(const double *A, const double *b, long long n, int D, int transpose, double *x,
                                cpfn_stream_t stream) {
  using namespace cpfn;
  if (n < 0 || D < 1 || D > 3) return CPFN_EINVAL;
  if (n == 0) return CPFN_OK;
  if (!A || !b || !x) return CPFN_EINVAL;
  const unsigned grid = static_cast<unsigned>((n + kLinThreads - 1) / kLinThreads);
  if (D == 1) small_solve_kernel<1><<<grid, kLinThreads, 0, as_stream(stream)>>>(A, b, n, transpose, x);
  else if (D == 2) small_solve_kernel<2><<<grid, kLinThreads, 0, as_stream(stream)>>>(A, b, n, transpose, x);
  else small_solve_kernel<3><<<grid, kLinThreads, 0, as_stream(stream)>>>(A, b, n, transpose, x);
  return check_launch();
}
