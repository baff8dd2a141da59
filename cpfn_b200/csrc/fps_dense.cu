// Farthest point sampling of ONE large cloud with the semantics of the reference's preprocessing
// (Preprocessing/preprocessing_sampling_lowres.py:14-42, numba on the CPU):
//
//   furthest_point_sampling(points, seeds, n)        min_dist = 1e6 everywhere, 0 at the seeds; n times:
//       take the arg-max (FIRST maximum), record it, min_dist = min(min_dist, ||p - p[index]||)
//   furthest_point_sampling_per_label(points, labels) one sample per distinct label: after every pick the
//       points of the picked label drop out (min_dist = 0); starts from a given index.
//
// These differ from the pointnet2 op (csrc/fps.cu): true distances (sqrt) instead of squared ones, first-index
// tie-break instead of the reduction-tree order, no ||p||^2 <= 1e-3 skip, seeds / labels.  The whole GPU works
// on the one cloud: a cooperative grid of one CTA per SM keeps every point and its running minimum in
// registers; per round a CTA publishes its best (distance, index) key with a round stamp and every CTA polls
// all stamps -- one L2 round trip per round, no separate grid barrier.  Distances use numpy's fp32 sequence
// sqrt((dx*dx + dy*dy) + dz*dz); the running minimum is float64 in the reference but only ever holds 1e6, 0 or
// float32 distances, all exact in float32.
#include "common.cuh"

namespace cpfn {
namespace {

constexpr int kDenseThreads = 1024;
constexpr int kDenseMaxCtas = 256;
constexpr int kDenseMaxSeeds = 1024;

// One 64-bit word per CTA and round parity carries the CTA's best candidate AND the round it belongs to:
//   [63:32] bits(min_dist)   [31:11] ~index (21 bits: N <= 2^21)   [10:0] round mod 2048
// so a single relaxed store publishes it and a single load both detects and reads it (no fence, one L2 trip).
// The buffer of a parity last held round r-2, so "round field == r" cannot be a stale match.
struct DenseWs {                                  // lives in the caller's workspace, zeroed by the launcher
  unsigned long long key[2][kDenseMaxCtas];
};
constexpr unsigned int kIdxMask = (1u << 21) - 1u;

__device__ __forceinline__ void st_relaxed_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ float np_dist(float ax, float ay, float az, float bx, float by, float bz) {
  const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
  return __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
}

// key = (bits(min_dist) << 32) | (~index & kIdxMask) : the maximum key is the largest distance, smallest index first
__device__ __forceinline__ unsigned long long warp_max_key(unsigned int hi, unsigned int lo) {
  const unsigned int M = __reduce_max_sync(0xffffffffu, hi);
  const unsigned int L = __reduce_max_sync(0xffffffffu, hi == M ? lo : 0u);
  return (static_cast<unsigned long long>(M) << 32) | L;
}

template <int PPT>
__global__ void __launch_bounds__(kDenseThreads, 1)
fps_dense_kernel(const float *__restrict__ pts, const int32_t *__restrict__ labels, const int32_t *__restrict__ seeds,
                 int n_seeds, int N, int n_out, int start_index, DenseWs *ws, int32_t *__restrict__ out) {
  __shared__ unsigned long long s_key[kDenseThreads / 32];
  __shared__ int s_seeds[kDenseMaxSeeds];
  __shared__ unsigned long long s_best;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int G = gridDim.x, cta = blockIdx.x;
  const int stride = G * kDenseThreads;
  const int first = cta * kDenseThreads + t;                       // point i of this thread: first + i * stride
  for (int i = t; i < n_seeds; i += kDenseThreads) s_seeds[i] = __ldg(seeds + i);
  __syncthreads();
  float px[PPT], py[PPT], pz[PPT], md[PPT];
  int lab[PPT];
#pragma unroll
  for (int i = 0; i < PPT; ++i) {
    const int k = first + i * stride;
    px[i] = py[i] = pz[i] = 0.f;
    md[i] = -1.f;                                                  // padding: below every real distance
    lab[i] = -1;
    if (k < N) {
      px[i] = __ldg(pts + 3ll * k); py[i] = __ldg(pts + 3ll * k + 1); pz[i] = __ldg(pts + 3ll * k + 2);
      md[i] = 1e6f;
      if (labels) lab[i] = __ldg(labels + k);
      for (int s = 0; s < n_seeds; ++s)
        if (s_seeds[s] == k) md[i] = 0.f;
    }
  }
  int index = start_index;                                         // per-label mode starts from a given point
  for (int round = 1; round <= n_out; ++round) {
    // ---- arg-max of the running minimum (skipped in the first per-label round: the start index is given) ----
    if (!(labels && round == 1)) {
      unsigned int hi = 0u, lo = 0u;
#pragma unroll
      for (int i = 0; i < PPT; ++i) {
        const unsigned int h = md[i] >= 0.f ? __float_as_uint(md[i]) : 0u;
        const unsigned int l = md[i] >= 0.f ? (~static_cast<unsigned int>(first + i * stride) & kIdxMask) : 0u;
        if (h > hi || (h == hi && l > lo)) { hi = h; lo = l; }
      }
      const unsigned long long wk = warp_max_key(hi, lo);
      if (lane == 0) s_key[warp] = wk;
      __syncthreads();
      if (warp == 0) {
        const unsigned long long v = s_key[lane];
        const unsigned long long ck = warp_max_key(static_cast<unsigned int>(v >> 32), static_cast<unsigned int>(v));
        if (lane == 0)
          st_relaxed_u64(&ws->key[round & 1][cta], (ck & 0xffffffff00000000ull) | ((ck & kIdxMask) << 11) |
                                                       static_cast<unsigned long long>(round & 2047));
      }
      // ---- every CTA gathers all CTAs' keys of this round ----
      unsigned int h2 = 0u, l2 = 0u;
      if (t < G) {
        unsigned long long v;
        do { v = ld_relaxed_u64(&ws->key[round & 1][t]); } while ((static_cast<unsigned int>(v) & 2047u) != static_cast<unsigned int>(round & 2047));
        h2 = static_cast<unsigned int>(v >> 32); l2 = (static_cast<unsigned int>(v) >> 11) & kIdxMask;
      }
      __syncthreads();                                             // s_key free again
      if (warp < (G + 31) / 32) {
        const unsigned long long gk = warp_max_key(h2, l2);
        if (lane == 0) s_key[warp] = gk;
      }
      __syncthreads();
      if (t == 0) {
        unsigned long long best = 0ull;
        for (int w = 0; w < (G + 31) / 32; ++w) best = s_key[w] > best ? s_key[w] : best;
        s_best = best;
      }
      __syncthreads();
      index = static_cast<int>(~static_cast<unsigned int>(s_best) & kIdxMask);
    }
    if (cta == 0 && t == 0) out[round - 1] = index;
    // ---- fold the new sample into the running minimum ----
    const float qx = __ldg(pts + 3ll * index), qy = __ldg(pts + 3ll * index + 1), qz = __ldg(pts + 3ll * index + 2);
    const int qlab = labels ? __ldg(labels + index) : 0;
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
      if (md[i] < 0.f) continue;
      md[i] = fminf(md[i], np_dist(px[i], py[i], pz[i], qx, qy, qz));
      if (labels && lab[i] == qlab) md[i] = 0.f;
    }
  }
}

}  // namespace
}  // namespace cpfn

using namespace cpfn;

extern "C" size_t cpfn_fps_dense_workspace_bytes(void) { return sizeof(DenseWs); }

extern "C" int cpfn_fps_dense(const float *points, int N, const int32_t *labels, const int32_t *seeds, int n_seeds,
                              int start_index, int n_out, int32_t *out, void *workspace, size_t workspace_bytes,
                              cpfn_stream_t stream) {
  if (!points || !out || N <= 0 || n_out <= 0 || n_seeds < 0 || n_seeds > kDenseMaxSeeds || (n_seeds > 0 && !seeds) ||
      start_index < 0 || start_index >= N || (labels && n_seeds > 0))
    return CPFN_EINVAL;
  if (!workspace || workspace_bytes < sizeof(DenseWs) || (reinterpret_cast<uintptr_t>(workspace) & 7)) return CPFN_EWORKSPACE;
  int sms = sm_count();
  if (sms <= 0) return CPFN_ELAUNCH;
  if (sms > kDenseMaxCtas) sms = kDenseMaxCtas;
  int G = (N + kDenseThreads - 1) / kDenseThreads;                 // one CTA per SM at most: all co-resident
  if (G > sms) G = sms;
  const long long per_thread = (static_cast<long long>(N) + static_cast<long long>(G) * kDenseThreads - 1) /
                               (static_cast<long long>(G) * kDenseThreads);
  if (per_thread > 8 || N > (1 << 21)) return CPFN_EINVAL;         // N <= 8 * 1024 * SMs (1.2 M points on a B200)
  cudaStream_t s = as_stream(stream);
  CPFN_CUDA_TRY(cudaMemsetAsync(workspace, 0, sizeof(DenseWs), s));
  DenseWs *ws = static_cast<DenseWs *>(workspace);
  void *args[] = {&points, &labels, &seeds, &n_seeds, &N, &n_out, &start_index, &ws, &out};
  const void *kern = per_thread <= 1   ? reinterpret_cast<const void *>(fps_dense_kernel<1>)
                     : per_thread <= 2 ? reinterpret_cast<const void *>(fps_dense_kernel<2>)
                     : per_thread <= 4 ? reinterpret_cast<const void *>(fps_dense_kernel<4>)
                                       : reinterpret_cast<const void *>(fps_dense_kernel<8>);
  // cooperative launch: the runtime refuses the launch unless all G CTAs can be resident at once, which the
  // stamp polling between CTAs relies on
  CPFN_CUDA_TRY(cudaLaunchCooperativeKernel(kern, dim3(G), dim3(kDenseThreads), args, 0, s));
  return check_launch();
}
