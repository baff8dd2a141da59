// Segmentation glue of the SPFN losses / metrics on the device:
//   hungarian_matching   SPFN/losses_implementation.py:11-30, SPFN/metric_implementation.py:9-30
//   compute_miou_loss    SPFN/losses_implementation.py:77-89 (forward sums; the backward is a gather, spfn/seg.py)
// The reference builds a one-hot W_gt [N, K'+1] per sample, multiplies it with W_pred (torch.mm), copies the K' x K
// cost matrix to the host and calls scipy.optimize.linear_sum_assignment there -- a device->host synchronisation and
// a Python loop over the batch in every training step.  Here:
//   cpfn_label_membership_sums   one pass over W [B,N,K]: S[b,g,k] = sum_{n: I[b,n]=g} W[b,n,k], the column sums, the
//                                label counts and n_gt[b] = max I[b,:] + 1 (fp32 per thread, fp64 across threads and
//                                CTAs, fixed order: deterministic);
//   cpfn_hungarian_matching      one warp per sample: the IoU cost matrix from those sums and the rectangular
//                                assignment of scipy's solver (shortest augmenting paths with dual variables, Crouse's
//                                variant of Jonker-Volgenant, in fp64), lane = column, INCLUDING its tie-breaking -- the
//                                scan order of the remaining columns and "prefer an unassigned column among equal
//                                minima" -- so degenerate cost matrices (all-zero rows) give scipy's answer too.
#include <math.h>

#include "common.cuh"

namespace cpfn {
namespace {

constexpr int kSegThreads = 256;
constexpr int kSegMaxK = 64;

// grid (chunks, B).  Thread t owns slot k = t % K and point lane t / K; a point lane adds w into ITS OWN row table
// s_tab[lane][g][k], so no two threads ever touch the same word.  Column K of a row counts the lane's points with that
// label (added by the lane's k = 0 thread; exact in fp32: a lane sees far fewer than 2^24 points).
template <typename IndexT>
__global__ void __launch_bounds__(kSegThreads)
seg_sums_kernel(const float *__restrict__ W, const IndexT *__restrict__ I, int N, int K, int G, int per_cta,
                double *__restrict__ part, int *__restrict__ n_gt) {
  extern __shared__ float s_tab[];                        // [lanes][G + 1][K + 1]  (row G: points without a label)
  const int b = blockIdx.y, t = threadIdx.x;
  const int K1 = K + 1;
  const int lanes = kSegThreads / K;
  const bool active = t < lanes * K;
  const int k = active ? t % K : 0, lane = t / K;
  for (int i = t; i < lanes * (G + 1) * K1; i += kSegThreads) s_tab[i] = 0.f;
  __syncthreads();
  const int n0 = blockIdx.x * per_cta, n1 = min(N, n0 + per_cta);
  int top = -1;
  if (active) {
    float *tab = s_tab + static_cast<size_t>(lane) * (G + 1) * K1 + k;
    float *cnt = s_tab + static_cast<size_t>(lane) * (G + 1) * K1 + K;
    const float *w = W + (static_cast<size_t>(b) * N) * K + k;
    const IndexT *lab = I + static_cast<size_t>(b) * N;
    for (int n = n0 + lane; n < n1; n += lanes) {
      const long long g = static_cast<long long>(lab[n]);
      top = max(top, static_cast<int>(g));
      const int row = (g >= 0 && g < G) ? static_cast<int>(g) : G;
      tab[row * K1] += __ldg(w + static_cast<size_t>(n) * K);
      if (k == 0) cnt[row * K1] += 1.f;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) top = max(top, __shfl_xor_sync(0xffffffffu, top, o));
  if ((t & 31) == 0 && top >= 0) atomicMax(n_gt + b, top + 1);
  __syncthreads();
  // fixed-order fp64 combination of the point lanes -> this CTA's partial [G + 1][K + 1]
  double *out = part + (static_cast<size_t>(b) * gridDim.x + blockIdx.x) * (G + 1) * K1;
  for (int i = t; i < (G + 1) * K1; i += kSegThreads) {
    double s = 0.0;
    for (int l = 0; l < lanes; ++l) s += static_cast<double>(s_tab[static_cast<size_t>(l) * (G + 1) * K1 + i]);
    out[i] = s;
  }
}

// S[b,g,k] (g < G), colsum[b,k] = sum over ALL points, count[b,g] = points with label g  (fp32 outputs): the chunks'
// partial tables added in a fixed order.  (Round 2: the label counts come from the tables too -- one CTA per cloud
// counting the labels again with shared-memory atomics took 2-3 ms on a 1 M-point cloud.)
__global__ void __launch_bounds__(kSegThreads)
seg_finish_kernel(const double *__restrict__ part, int K, int G, int chunks,
                  float *__restrict__ S, float *__restrict__ colsum, float *__restrict__ count) {
  const int b = blockIdx.x, t = threadIdx.x;
  const int K1 = K + 1;
  extern __shared__ double s_tot[];                       // [(G + 1)][K + 1]
  const double *pb = part + static_cast<size_t>(b) * chunks * (G + 1) * K1;
  for (int i = t; i < (G + 1) * K1; i += kSegThreads) {
    double s = 0.0;
    for (int c = 0; c < chunks; ++c) s += pb[static_cast<size_t>(c) * (G + 1) * K1 + i];
    s_tot[i] = s;
    const int g = i / K1, kk = i - g * K1;
    if (g < G) {
      if (kk < K) S[(static_cast<size_t>(b) * G + g) * K + kk] = static_cast<float>(s);
      else count[static_cast<size_t>(b) * G + g] = static_cast<float>(s);
    }
  }
  __syncthreads();
  for (int kk = t; kk < K; kk += kSegThreads) {
    double s = 0.0;
    for (int g = 0; g <= G; ++g) s += s_tot[g * K1 + kk];
    colsum[static_cast<size_t>(b) * K + kk] = static_cast<float>(s);
  }
}

// One warp per sample; lane j = column j (K <= 32).  scipy/optimize/rectangular_lsap (maximize=True => cost negated).
__global__ void __launch_bounds__(128)
hungarian_kernel(const float *__restrict__ S, const float *__restrict__ colsum, const float *__restrict__ count,
                 const int *__restrict__ n_gt, int B, int K, int G, long long *__restrict__ matching,
                 unsigned char *__restrict__ mask) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * 4 + warp;
  if (b >= B) return;
  __shared__ double s_cost[4][32 * 32];      // [row][col], negated IoU
  __shared__ double s_u[4][32], s_spc[4][32];
  __shared__ int s_col4row[4][32], s_row4col[4][32], s_path[4][32], s_remaining[4][32];
  __shared__ unsigned char s_SR[4][32], s_SC[4][32];
  double *cost = s_cost[warp], *u = s_u[warp], *spc = s_spc[warp];
  int *col4row = s_col4row[warp], *row4col = s_row4col[warp], *path = s_path[warp], *remaining = s_remaining[warp];
  unsigned char *SR = s_SR[warp], *SC = s_SC[warp];
  const int nr = min(min(n_gt[b], G), K), nc = K;   // K' <= K rows (gt labels), K columns (predicted slots)
  // cost = dot / clamp(count_g + colsum_k - dot, 1e-10) in fp32 as the reference computes it, negated in fp64
  for (int g = 0; g < nr; ++g)
    if (lane < nc) {
      const float dot = S[(static_cast<size_t>(b) * G + g) * K + lane];
      const float den = fmaxf(__fsub_rn(__fadd_rn(count[static_cast<size_t>(b) * G + g], colsum[static_cast<size_t>(b) * K + lane]), dot), 1e-10f);
      cost[g * 32 + lane] = -static_cast<double>(__fdiv_rn(dot, den));
    }
  double v = 0.0;                                   // dual of this lane's column
  if (lane < 32) { u[lane] = 0.0; col4row[lane] = -1; row4col[lane] = -1; }
  __syncwarp();
  for (int cur = 0; cur < nr; ++cur) {
    double min_val = 0.0;
    int i = cur, num_remaining = nc, sink = -1;
    if (lane < nc) remaining[lane] = nc - lane - 1;  // scipy fills the list in reverse order
    SR[lane] = 0; SC[lane] = 0; spc[lane] = INFINITY;
    __syncwarp();
    while (sink == -1) {
      if (lane == 0) SR[i] = 1;
      // lane `it` handles remaining[it]
      const bool have = lane < num_remaining;
      const int j = have ? remaining[lane] : 0;
      double cand = INFINITY;
      bool unassigned = false;
      // the column duals live in registers of lane == column: fetch v[j] for this lane's j (all lanes shuffle)
      const double vj = __shfl_sync(0xffffffffu, v, have ? j : 0);
      if (have) {
        const double r = min_val + cost[i * 32 + j] - u[i] - vj;
        if (r < spc[j]) { path[j] = i; spc[j] = r; }
        cand = spc[j];
        unassigned = row4col[j] == -1;
      }
      __syncwarp();
      // lowest = min over the list; index = last unassigned entry among the minima if there is one, else the first
      double lowest = cand;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) lowest = fmin(lowest, __shfl_xor_sync(0xffffffffu, lowest, o));
      const unsigned int at_min = __ballot_sync(0xffffffffu, have && cand == lowest);
      const unsigned int at_min_free = __ballot_sync(0xffffffffu, have && cand == lowest && unassigned);
      if (lowest == INFINITY || at_min == 0u) { sink = -2; break; }          // infeasible (cannot happen: finite costs)
      // scipy scans the list in order: the first minimum sets the index, any LATER entry with the same value
      // replaces it when its column is unassigned
      const int first = __ffs(at_min) - 1;
      const unsigned int later_free = first == 31 ? 0u : (at_min_free & ~((2u << first) - 1u));
      const int index = later_free ? 31 - __clz(later_free) : first;
      min_val = lowest;
      const int jsel = __shfl_sync(0xffffffffu, j, index);
      const int owner = row4col[jsel];
      if (owner == -1) sink = jsel; else i = owner;
      __syncwarp();
      if (lane == 0) {
        SC[jsel] = 1;
        remaining[index] = remaining[num_remaining - 1];
      }
      --num_remaining;
      __syncwarp();
    }
    if (sink < 0) break;
    // dual update
    if (lane == 0) u[cur] += min_val;
    __syncwarp();
    if (lane < nr && SR[lane] && lane != cur) u[lane] += min_val - spc[col4row[lane]];
    if (lane < nc && SC[lane]) v -= min_val - spc[lane];
    __syncwarp();
    // augment along the path (sequential, lane 0)
    if (lane == 0) {
      int j = sink;
      while (true) {
        const int ii = path[j];
        row4col[j] = ii;
        const int prev = col4row[ii];
        col4row[ii] = j;
        j = prev;
        if (ii == cur) break;
      }
    }
    __syncwarp();
  }
  if (lane < K) {
    matching[static_cast<size_t>(b) * K + lane] = lane < nr ? col4row[lane] : 0;
    if (mask) mask[static_cast<size_t>(b) * K + lane] = lane < nr ? 1 : 0;
  }
}

}  // namespace
}  // namespace cpfn

using namespace cpfn;

extern "C" size_t cpfn_seg_workspace_bytes(int B, int N, int K, int G) {
  if (B <= 0 || N <= 0 || K <= 0 || G <= 0 || K > kSegMaxK || G > kSegMaxK) return 0;
  const int sms = sm_count() > 0 ? sm_count() : 148;
  int chunks = (2 * sms + B - 1) / B;
  if (chunks > (N + 255) / 256) chunks = (N + 255) / 256;
  if (chunks < 1) chunks = 1;
  return sizeof(double) * static_cast<size_t>(B) * chunks * (G + 1) * (K + 1) + 256;
}

extern "C" int cpfn_label_membership_sums(const float *W, const void *labels, int labels_are_int64, int B, int N, int K,
                                          int G, float *S, float *colsum, float *count, int32_t *n_gt, void *workspace,
                                          size_t workspace_bytes, cpfn_stream_t stream) {
  if (B < 0 || N < 0 || K <= 0 || G <= 0 || K > kSegMaxK || G > kSegMaxK) return CPFN_EINVAL;
  if (B == 0) return CPFN_OK;
  if (!W || !labels || !S || !colsum || !count || !n_gt || N == 0 || B > 65535) return CPFN_EINVAL;
  if (!workspace || workspace_bytes < cpfn_seg_workspace_bytes(B, N, K, G)) return CPFN_EWORKSPACE;
  cudaStream_t st = as_stream(stream);
  const int sms = sm_count() > 0 ? sm_count() : 148;
  int chunks = (2 * sms + B - 1) / B;
  if (chunks > (N + 255) / 256) chunks = (N + 255) / 256;
  if (chunks < 1) chunks = 1;
  const int per_cta = (N + chunks - 1) / chunks;
  const int lanes = kSegThreads / K;
  const size_t smem = sizeof(float) * static_cast<size_t>(lanes) * (G + 1) * (K + 1);
  const size_t fsmem = sizeof(double) * static_cast<size_t>(G + 1) * (K + 1);
  double *part = static_cast<double *>(workspace);
  CPFN_CUDA_TRY(cudaMemsetAsync(n_gt, 0, sizeof(int32_t) * B, st));
  if (labels_are_int64) {
    if (smem > 48 * 1024) CPFN_CUDA_TRY(cudaFuncSetAttribute(seg_sums_kernel<long long>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    seg_sums_kernel<long long><<<dim3(chunks, B), kSegThreads, smem, st>>>(W, static_cast<const long long *>(labels), N, K, G, per_cta, part, n_gt);
    seg_finish_kernel<<<B, kSegThreads, fsmem, st>>>(part, K, G, chunks, S, colsum, count);
  } else {
    if (smem > 48 * 1024) CPFN_CUDA_TRY(cudaFuncSetAttribute(seg_sums_kernel<int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    seg_sums_kernel<int32_t><<<dim3(chunks, B), kSegThreads, smem, st>>>(W, static_cast<const int32_t *>(labels), N, K, G, per_cta, part, n_gt);
    seg_finish_kernel<<<B, kSegThreads, fsmem, st>>>(part, K, G, chunks, S, colsum, count);
  }
  return check_launch();
}

extern "C" int cpfn_hungarian_matching(const float *S, const float *colsum, const float *count, const int32_t *n_gt, int B,
                                       int K, int G, long long *matching, unsigned char *mask, cpfn_stream_t stream) {
  if (B < 0 || K <= 0 || K > 32 || G <= 0 || G > kSegMaxK) return CPFN_EINVAL;
  if (B == 0) return CPFN_OK;
  if (!S || !colsum || !count || !n_gt || !matching) return CPFN_EINVAL;
  hungarian_kernel<<<(B + 3) / 4, 128, 0, as_stream(stream)>>>(S, colsum, count, n_gt, B, K, G, matching, mask);
  return check_launch();
}
