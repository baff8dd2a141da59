// three_nn, three_weighted_sum (+grad), gather_points (+grad), group_points
// (+grad) for sm_100a.
//
// Semantics follow PointNet2/pointnet2_ops/cuda_ops/src/interpolate_gpu.cu
// (three_nn :9-59, three_weighted_sum :72-101, grad :116-143),
// src/sampling_gpu.cu (gather :8-53) and src/group_points_gpu.cu (:8-74).
//
// Design notes (vs the reference's one block per cloud):
//   * three_nn: one thread per unknown point, the known cloud staged through
//     shared memory as float4 (one broadcast LDS.128 per candidate), grid over
//     (unknown tiles x clouds) so every SM has work.  The reference keeps its
//     three best distances as doubles initialised to 1e40; float -> double is
//     exact and monotonic, so the same decisions are taken here in fp32 with
//     +inf as the initial value ((float)1e40 == +inf is also what the reference
//     stores when fewer than three candidates exist).
//   * weighted sum / gather / group: one thread per output column j (coalesced
//     stores), looping over a tile of channels; 64-bit offsets throughout.
//   * gradients: scatter-add with fire-and-forget fp32 RED (atomicAdd without
//     a return value), destination zeroed on the same stream first.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "nn_grid.cuh"

namespace cpfn {
namespace {

constexpr int kNnThreads = 256;
constexpr int kNnTile = 2048;  // known points per shared tile (32 KB as float4)

__global__ void __launch_bounds__(kNnThreads)
three_nn_kernel(const float *__restrict__ unknown, const float *__restrict__ known, int n, int m,
                float *__restrict__ dist2, int32_t *__restrict__ idx) {
  extern __shared__ float4 s_known[];      // min(m, kNnTile) records: small enough to co-reside with the MLP chains
  const int b = blockIdx.y;
  const int j = blockIdx.x * kNnThreads + threadIdx.x;
  const float *kn = known + static_cast<size_t>(b) * m * 3;
  float ux = 0.f, uy = 0.f, uz = 0.f;
  if (j < n) {
    const float *u = unknown + (static_cast<size_t>(b) * n + j) * 3;
    ux = __ldg(u); uy = __ldg(u + 1); uz = __ldg(u + 2);
  }
  float b1 = INFINITY, b2 = INFINITY, b3 = INFINITY;
  int i1 = 0, i2 = 0, i3 = 0;
  for (int t0 = 0; t0 < m; t0 += kNnTile) {
    const int tn = min(kNnTile, m - t0);
    __syncthreads();
    for (int k = threadIdx.x; k < tn; k += kNnThreads) {
      const float *s = kn + static_cast<size_t>(t0 + k) * 3;
      s_known[k] = make_float4(__ldg(s), __ldg(s + 1), __ldg(s + 2), 0.f);
    }
    __syncthreads();
#pragma unroll 4
    for (int k = 0; k < tn; ++k) {
      const float4 c = s_known[k];
      const float d = sqdist3(ux, uy, uz, c.x, c.y, c.z);
      if (d < b3) {  // b1 <= b2 <= b3, so this guards the reference's whole cascade
        const int kk = t0 + k;
        if (d < b1) {
          b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = kk;
        } else if (d < b2) {
          b3 = b2; i3 = i2; b2 = d; i2 = kk;
        } else {
          b3 = d; i3 = kk;
        }
      }
    }
  }
  if (j < n) {
    const size_t o = (static_cast<size_t>(b) * n + j) * 3;
    dist2[o] = b1; dist2[o + 1] = b2; dist2[o + 2] = b3;
    idx[o] = i1; idx[o + 1] = i2; idx[o + 2] = i3;
  }
}

// Grid form (nn_grid.cuh) for known clouds of <= kNnGridMax points: same indices and distances, ~10x fewer
// candidates per query.
constexpr int kNnGridQ = 2;

__global__ void __launch_bounds__(kNnThreads)
three_nn_grid_kernel(const float *__restrict__ unknown, const float *__restrict__ known, int n, int m,
                     float *__restrict__ dist2, int32_t *__restrict__ idx) {
  extern __shared__ __align__(16) unsigned char s_grid_raw[];
  const int b = blockIdx.y;
  NnGrid g;
  float4 *recs;
  int *cell_start;
  nn_grid_build(known + static_cast<size_t>(b) * m * 3, m, s_grid_raw, g, recs, cell_start);
#pragma unroll 1
  for (int q = 0; q < kNnGridQ; ++q) {
    const int j = (blockIdx.x * kNnGridQ + q) * kNnThreads + threadIdx.x;
    if (j >= n) continue;
    const float *u = unknown + (static_cast<size_t>(b) * n + j) * 3;
    const Nn3 r = nn_grid_query(__ldg(u), __ldg(u + 1), __ldg(u + 2), g, recs, cell_start);
    const size_t o = (static_cast<size_t>(b) * n + j) * 3;
    dist2[o] = r.d1; dist2[o + 1] = r.d2; dist2[o + 2] = r.d3;
    idx[o] = r.i1; idx[o + 1] = r.i2; idx[o + 2] = r.i3;
  }
}

constexpr int kColThreads = 256;
constexpr int kChanTile = 16;

// out[b,c,j] = fma(p3,w3, fma(p1,w1, p2*w2))   (interpolate_gpu.cu:98-99 as nvcc contracts it)
__global__ void __launch_bounds__(kColThreads)
three_weighted_sum_kernel(const float *__restrict__ points, const int32_t *__restrict__ idx,
                          const float *__restrict__ weight, int C, int M, int n,
                          float *__restrict__ out) {
  const int b = blockIdx.z;
  const int j = blockIdx.x * kColThreads + threadIdx.x;
  if (j >= n) return;
  const size_t o3 = (static_cast<size_t>(b) * n + j) * 3;
  const int i1 = __ldg(idx + o3), i2 = __ldg(idx + o3 + 1), i3 = __ldg(idx + o3 + 2);
  const float w1 = __ldg(weight + o3), w2 = __ldg(weight + o3 + 1), w3 = __ldg(weight + o3 + 2);
  const int c0 = blockIdx.y * kChanTile, c1 = min(C, c0 + kChanTile);
  for (int c = c0; c < c1; ++c) {
    const float *row = points + (static_cast<size_t>(b) * C + c) * M;
    const float v = __fmaf_rn(__ldg(row + i3), w3, __fmaf_rn(__ldg(row + i1), w1,
                                                           __fmul_rn(__ldg(row + i2), w2)));
    out[(static_cast<size_t>(b) * C + c) * n + j] = v;
  }
}

__global__ void __launch_bounds__(kColThreads)
three_weighted_sum_grad_kernel(const float *__restrict__ grad_out,
                               const int32_t *__restrict__ idx,
                               const float *__restrict__ weight, int C, int n, int M,
                               float *__restrict__ grad_points) {
  const int b = blockIdx.z;
  const int j = blockIdx.x * kColThreads + threadIdx.x;
  if (j >= n) return;
  const size_t o3 = (static_cast<size_t>(b) * n + j) * 3;
  const int i1 = __ldg(idx + o3), i2 = __ldg(idx + o3 + 1), i3 = __ldg(idx + o3 + 2);
  const float w1 = __ldg(weight + o3), w2 = __ldg(weight + o3 + 1), w3 = __ldg(weight + o3 + 2);
  const int c0 = blockIdx.y * kChanTile, c1 = min(C, c0 + kChanTile);
  for (int c = c0; c < c1; ++c) {
    const float g = __ldg(grad_out + (static_cast<size_t>(b) * C + c) * n + j);
    float *row = grad_points + (static_cast<size_t>(b) * C + c) * M;
    atomicAdd(row + i1, __fmul_rn(g, w1));
    atomicAdd(row + i2, __fmul_rn(g, w2));
    atomicAdd(row + i3, __fmul_rn(g, w3));
  }
}

// out[b,c,j] = points[b,c,idx[b,j]]; group_points is the same map with
// j running over S*K.
__global__ void __launch_bounds__(kColThreads)
gather_kernel(const float *__restrict__ points, const int32_t *__restrict__ idx, int C, int N,
              long long M, float *__restrict__ out) {
  const int b = blockIdx.z;
  const long long j = static_cast<long long>(blockIdx.x) * kColThreads + threadIdx.x;
  if (j >= M) return;
  const int a = __ldg(idx + static_cast<size_t>(b) * M + j);
  const int c0 = blockIdx.y * kChanTile, c1 = min(C, c0 + kChanTile);
  for (int c = c0; c < c1; ++c)
    out[(static_cast<size_t>(b) * C + c) * M + j] =
        __ldg(points + (static_cast<size_t>(b) * C + c) * N + a);
}

__global__ void __launch_bounds__(kColThreads)
gather_grad_kernel(const float *__restrict__ grad_out, const int32_t *__restrict__ idx, int C,
                   int N, long long M, float *__restrict__ grad_points) {
  const int b = blockIdx.z;
  const long long j = static_cast<long long>(blockIdx.x) * kColThreads + threadIdx.x;
  if (j >= M) return;
  const int a = __ldg(idx + static_cast<size_t>(b) * M + j);
  const int c0 = blockIdx.y * kChanTile, c1 = min(C, c0 + kChanTile);
  for (int c = c0; c < c1; ++c)
    atomicAdd(grad_points + (static_cast<size_t>(b) * C + c) * N + a,
              __ldg(grad_out + (static_cast<size_t>(b) * C + c) * M + j));
}

inline bool grid_ok(long long cols, int C, int B) {
  return (cols + kColThreads - 1) / kColThreads <= 2147483647LL &&
         (C + kChanTile - 1) / kChanTile <= 65535 && B <= 65535;
}

int gather_impl(const float *points, const int32_t *idx, int B, int C, int N, long long M,
                float *out, cudaStream_t st) {
  if (B < 0 || C < 0 || N < 0 || M < 0) return CPFN_EINVAL;
  if (B == 0 || C == 0 || M == 0) return CPFN_OK;
  if (!points || !idx || !out || N == 0 || !grid_ok(M, C, B)) return CPFN_EINVAL;
  dim3 grid(static_cast<unsigned>((M + kColThreads - 1) / kColThreads),
            (C + kChanTile - 1) / kChanTile, B);
  gather_kernel<<<grid, kColThreads, 0, st>>>(points, idx, C, N, M, out);
  return check_launch();
}

int gather_grad_impl(const float *grad_out, const int32_t *idx, int B, int C, int N, long long M,
                     float *grad_points, cudaStream_t st) {
  if (B < 0 || C < 0 || N < 0 || M < 0) return CPFN_EINVAL;
  if (B == 0 || C == 0 || N == 0) return CPFN_OK;
  if (!grad_points) return CPFN_EINVAL;
  CPFN_CUDA_TRY(cudaMemsetAsync(grad_points, 0, sizeof(float) * size_t(B) * C * N, st));
  if (M == 0) return CPFN_OK;
  if (!grad_out || !idx || !grid_ok(M, C, B)) return CPFN_EINVAL;
  dim3 grid(static_cast<unsigned>((M + kColThreads - 1) / kColThreads),
            (C + kChanTile - 1) / kChanTile, B);
  gather_grad_kernel<<<grid, kColThreads, 0, st>>>(grad_out, idx, C, N, M, grad_points);
  return check_launch();
}

}  // namespace
}  // namespace cpfn

extern "C" int cpfn_three_nn(const float *unknown, const float *known, int B, int n, int m,
                             float *dist2, int32_t *idx, cpfn_stream_t stream) {
  using namespace cpfn;
  if (B < 0 || n < 0 || m < 0) return CPFN_EINVAL;
  if (B == 0 || n == 0) return CPFN_OK;
  if (!unknown || !dist2 || !idx || (m > 0 && !known) || B > 65535) return CPFN_EINVAL;
  if (m >= kNnGridMin && m <= kNnGridMax && getenv("CPFN_NN_NO_GRID") == nullptr) {
    const size_t gsmem = nn_grid_smem_bytes(m);
    if (gsmem > 48 * 1024)
      CPFN_CUDA_TRY(cudaFuncSetAttribute(three_nn_grid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(gsmem)));
    dim3 ggrid((n + kNnThreads * kNnGridQ - 1) / (kNnThreads * kNnGridQ), B);
    three_nn_grid_kernel<<<ggrid, kNnThreads, gsmem, as_stream(stream)>>>(unknown, known, n, m, dist2, idx);
    return check_launch();
  }
  dim3 grid((n + kNnThreads - 1) / kNnThreads, B);
  const size_t smem = sizeof(float4) * static_cast<size_t>(m < kNnTile ? (m > 0 ? m : 1) : kNnTile);
  three_nn_kernel<<<grid, kNnThreads, smem, as_stream(stream)>>>(unknown, known, n, m, dist2, idx);
  return check_launch();
}

extern "C" int cpfn_three_weighted_sum(const float *points, const int32_t *idx,
                                       const float *weight, int B, int C, int M, int n,
                                       float *out, cpfn_stream_t stream) {
  using namespace cpfn;
  if (B < 0 || C < 0 || M < 0 || n < 0) return CPFN_EINVAL;
  if (B == 0 || C == 0 || n == 0) return CPFN_OK;
  if (!points || !idx || !weight || !out || M == 0 || !grid_ok(n, C, B)) return CPFN_EINVAL;
  dim3 grid((n + kColThreads - 1) / kColThreads, (C + kChanTile - 1) / kChanTile, B);
  three_weighted_sum_kernel<<<grid, kColThreads, 0, as_stream(stream)>>>(points, idx, weight, C,
                                                                         M, n, out);
  return check_launch();
}

extern "C" int cpfn_three_weighted_sum_grad(const float *grad_out, const int32_t *idx,
                                            const float *weight, int B, int C, int n, int M,
                                            float *grad_points, cpfn_stream_t stream) {
  using namespace cpfn;
  if (B < 0 || C < 0 || M < 0 || n < 0) return CPFN_EINVAL;
  if (B == 0 || C == 0 || M == 0) return CPFN_OK;
  if (!grad_points) return CPFN_EINVAL;
  cudaStream_t st = as_stream(stream);
  CPFN_CUDA_TRY(cudaMemsetAsync(grad_points, 0, sizeof(float) * size_t(B) * C * M, st));
  if (n == 0) return CPFN_OK;
  if (!grad_out || !idx || !weight || !grid_ok(n, C, B)) return CPFN_EINVAL;
  dim3 grid((n + kColThreads - 1) / kColThreads, (C + kChanTile - 1) / kChanTile, B);
  three_weighted_sum_grad_kernel<<<grid, kColThreads, 0, st>>>(grad_out, idx, weight, C, n, M,
                                                               grad_points);
  return check_launch();
}

extern "C" int cpfn_gather_points(const float *points, const int32_t *idx, int B, int C, int N,
                                  int M, float *out, cpfn_stream_t stream) {
  return cpfn::gather_impl(points, idx, B, C, N, M, out, cpfn::as_stream(stream));
}

extern "C" int cpfn_gather_points_grad(const float *grad_out, const int32_t *idx, int B, int C,
                                       int N, int M, float *grad_points, cpfn_stream_t stream) {
  return cpfn::gather_grad_impl(grad_out, idx, B, C, N, M, grad_points, cpfn::as_stream(stream));
}

extern "C" int cpfn_group_points(const float *points, const int32_t *idx, int B, int C, int N,
                                 int S, int K, float *out, cpfn_stream_t stream) {
  if (S < 0 || K < 0) return CPFN_EINVAL;
  return cpfn::gather_impl(points, idx, B, C, N, static_cast<long long>(S) * K, out,
                           cpfn::as_stream(stream));
}

extern "C" int cpfn_group_points_grad(const float *grad_out, const int32_t *idx, int B, int C,
                                      int N, int S, int K, float *grad_points,
                                      cpfn_stream_t stream) {
  if (S < 0 || K < 0) return CPFN_EINVAL;
  return cpfn::gather_grad_impl(grad_out, idx, B, C, N, static_cast<long long>(S) * K,
                                grad_points, cpfn::as_stream(stream));
}
