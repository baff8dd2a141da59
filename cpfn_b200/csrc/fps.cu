// Furthest point sampling for sm_100a.
//
// Semantics follow the reference kernel bit for bit
// (PointNet2/pointnet2_ops/cuda_ops/src/sampling_gpu.cu:63-159, host wrapper
// src/sampling.cpp:65-86); see SURVEY.md appendix A.1:
//   * idx[0] = 0; running minimum `temp` starts at 1e10;
//   * points with |p|^2 <= 1e-3 (double compare) are never updated and never
//     selected;
//   * d = fma(dz,dz,fma(dx,dx,dy*dy)), d2 = fminf(d, temp), argmax with strict
//     '>' over the reference's block of T = opt_n_threads(N) threads: among
//     equal maxima the winner is the reference thread with the smallest
//     bit-reversed id, then the smallest k inside that thread.
//
// Design (not the reference's): one CTA per cloud keeps every point AND its
// running minimum in registers for the whole kernel (no re-read of xyz/temp
// per round -- the reference streams 20 B/point/round through L2).  A round is
//   PPT x (3 FADD, FMUL, 2 FFMA, FMNMX, FSETP, 2 SEL)  per thread,
//   two REDUX.MAX per warp, one 8-byte STS per warp, ONE __syncthreads,
//   two REDUX.MAX over the per-warp keys (done redundantly by every warp so no
//   second barrier is needed), one broadcast LDS.x3 of the new centroid.
// The argmax key is 64 bits: hi = bits(d2) (non-negative floats order like
// uint32), lo = ~rank where rank encodes the reference tie-break order.
#include <math.h>

#include "common.cuh"

namespace cpfn {
namespace {

constexpr int kMaxSmemN = 16384;  // coords in shared memory (SoA), temp in registers

__host__ __device__ __forceinline__ uint32_t fps_rank(uint32_t k, int log2T) {
#ifdef __CUDA_ARCH__
  const uint32_t r = log2T ? (__brev(k & ((1u << log2T) - 1u)) >> (32 - log2T)) : 0u;
#else
  uint32_t r = 0, v = k & ((1u << log2T) - 1u);
  for (int i = 0; i < log2T; ++i) r |= ((v >> i) & 1u) << (log2T - 1 - i);
#endif
  return (r << 20) | (k >> log2T);
}

__device__ __forceinline__ uint32_t fps_unrank(uint32_t rank, int log2T) {
  const uint32_t r = rank >> 20, q = rank & 0xFFFFFu;
  const uint32_t tref = log2T ? (__brev(r) >> (32 - log2T)) : 0u;
  return (q << log2T) + tref;
}

// Block-wide argmax of (hi, lo) keys.  Every thread returns the winning lo
// (0 when no thread holds a valid candidate).  `slot` is a double-buffered
// [2][32] array; `buf` alternates per call.  One __syncthreads per call.
__device__ __forceinline__ uint32_t block_argmax(uint32_t hi, uint32_t lo,
                                                 unsigned long long (*slot)[32],
                                                 int buf, int nwarps) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t M = __reduce_max_sync(0xffffffffu, hi);
  uint32_t L = __reduce_max_sync(0xffffffffu, hi == M ? lo : 0u);
  if (nwarps == 1) return L;
  if (lane == 0) slot[buf][warp] = (static_cast<unsigned long long>(M) << 32) | L;
  __syncthreads();
  const unsigned long long v = lane < nwarps ? slot[buf][lane] : 0ull;
  const uint32_t h2 = static_cast<uint32_t>(v >> 32), l2 = static_cast<uint32_t>(v);
  M = __reduce_max_sync(0xffffffffu, h2);
  L = __reduce_max_sync(0xffffffffu, h2 == M ? l2 : 0u);
  return L;
}

// Stage xyz [N,3] (AoS, global) into shared SoA sx|sy|sz with pitch Np.
__device__ __forceinline__ void stage_soa(const float *__restrict__ p, int N, int Np,
                                          float *s) {
  for (int f = threadIdx.x; f < 3 * N; f += blockDim.x) {
    const int k = f / 3, c = f - 3 * k;
    s[c * Np + k] = __ldg(p + f);
  }
}

// PPT points per thread, thread t owns k = t + i*blockDim.x.  blockDim.x is a
// multiple of T (or PPT == 1), so scanning i upward visits a thread's points
// in increasing reference rank and the strict '>' keeps the right one on ties.
template <int PPT, bool COORDS_IN_REGS>
__global__ void __launch_bounds__(1024, 1)
fps_cta_kernel(const float *__restrict__ xyz, int N, int m, int log2T,
               int32_t *__restrict__ idx) {
  extern __shared__ float s_xyz[];
  __shared__ unsigned long long slot[2][32];
  const int NT = blockDim.x, t = threadIdx.x, nwarps = NT >> 5;
  const int Np = (N + 31) & ~31;
  const float *p = xyz + static_cast<size_t>(blockIdx.x) * N * 3;
  int32_t *out = idx + static_cast<size_t>(blockIdx.x) * m;
  float *sx = s_xyz, *sy = s_xyz + Np, *sz = s_xyz + 2 * Np;

  stage_soa(p, N, Np, s_xyz);
  __syncthreads();

  float px[PPT], py[PPT], pz[PPT], tmp[PPT];
#pragma unroll
  for (int i = 0; i < PPT; ++i) {
    const int k = t + i * NT;
    float x = 0.f, y = 0.f, z = 0.f;
    bool live = false;
    if (k < N) {
      x = sx[k]; y = sy[k]; z = sz[k];
      live = !(static_cast<double>(sqnorm3(x, y, z)) <= 1e-3);
    }
    if (COORDS_IN_REGS) { px[i] = x; py[i] = y; pz[i] = z; }
    // -1 marks "never a candidate": fminf(d, -1) stays -1 and -1 > best(-1) is false.
    tmp[i] = live ? 1e10f : -1.0f;
  }

  if (t == 0) out[0] = 0;
  float cx = sx[0], cy = sy[0], cz = sz[0];
  for (int j = 1; j < m; ++j) {
    float best = -1.0f;
    int bi = 0;
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
      float x, y, z;
      if (COORDS_IN_REGS) { x = px[i]; y = py[i]; z = pz[i]; }
      else { const int k = min(t + i * NT, Np - 1); x = sx[k]; y = sy[k]; z = sz[k]; }
      const float d2 = fminf(sqdist3(x, y, z, cx, cy, cz), tmp[i]);
      tmp[i] = d2;
      if (d2 > best) { best = d2; bi = i; }
    }
    uint32_t hi = 0u, lo = 0u;
    if (best >= 0.0f) {
      hi = __float_as_uint(best);
      lo = ~fps_rank(static_cast<uint32_t>(t + bi * NT), log2T);
    }
    const uint32_t L = block_argmax(hi, lo, slot, j & 1, nwarps);
    const int old = L ? static_cast<int>(fps_unrank(~L, log2T)) : 0;
    if (t == 0) out[j] = old;
    cx = sx[old]; cy = sy[old]; cz = sz[old];
  }
}

// Any N: running minimum in a global workspace [B,N], coordinates re-read
// through L1/L2 every round.  Only used beyond kMaxSmemN points per cloud.
__global__ void __launch_bounds__(1024, 1)
fps_stream_kernel(const float *__restrict__ xyz, int N, int m, int log2T,
                  float *__restrict__ temp, int32_t *__restrict__ idx) {
  __shared__ unsigned long long slot[2][32];
  const int NT = blockDim.x, t = threadIdx.x, nwarps = NT >> 5;
  const float *p = xyz + static_cast<size_t>(blockIdx.x) * N * 3;
  float *tmp = temp + static_cast<size_t>(blockIdx.x) * N;
  int32_t *out = idx + static_cast<size_t>(blockIdx.x) * m;
  for (int k = t; k < N; k += NT) {
    const float x = p[3 * k], y = p[3 * k + 1], z = p[3 * k + 2];
    tmp[k] = (static_cast<double>(sqnorm3(x, y, z)) <= 1e-3) ? -1.0f : 1e10f;
  }
  if (t == 0) out[0] = 0;
  float cx = p[0], cy = p[1], cz = p[2];
  for (int j = 1; j < m; ++j) {
    float best = -1.0f;
    int bk = 0;
    for (int k = t; k < N; k += NT) {
      const float d2 = fminf(sqdist3(p[3 * k], p[3 * k + 1], p[3 * k + 2], cx, cy, cz), tmp[k]);
      tmp[k] = d2;
      if (d2 > best) { best = d2; bk = k; }
    }
    uint32_t hi = 0u, lo = 0u;
    if (best >= 0.0f) {
      hi = __float_as_uint(best);
      lo = ~fps_rank(static_cast<uint32_t>(bk), log2T);
    }
    const uint32_t L = block_argmax(hi, lo, slot, j & 1, nwarps);
    const int old = L ? static_cast<int>(fps_unrank(~L, log2T)) : 0;
    if (t == 0) out[j] = old;
    cx = p[3 * old]; cy = p[3 * old + 1]; cz = p[3 * old + 2];
  }
}

// include/cuda_utils.h:15-19 of the reference, restated: the reference picks
// its block size (hence its tie-break order) through double log()/log(2.0).
int ref_log2_threads(int n) {
  int pow_2 = static_cast<int>(log(static_cast<double>(n)) / log(2.0));
  if (pow_2 > 9) pow_2 = 9;
  if (pow_2 < 0) pow_2 = 0;
  return pow_2;
}

template <int PPT, bool REGS>
int launch_cta(const float *xyz, int B, int N, int m, int log2T, int NT, int32_t *idx,
               cudaStream_t st) {
  const size_t smem = 3u * static_cast<size_t>((N + 31) & ~31) * sizeof(float);
  auto kern = fps_cta_kernel<PPT, REGS>;
  if (smem + 1024 > 48 * 1024)  // static shared memory counts against the 48 KB default
    CPFN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem)));
  kern<<<B, NT, smem, st>>>(xyz, N, m, log2T, idx);
  return check_launch();
}

}  // namespace
}  // namespace cpfn

extern "C" size_t cpfn_fps_workspace_bytes(int B, int N) {
  if (B <= 0 || N <= cpfn::kMaxSmemN) return 0;
  return static_cast<size_t>(B) * static_cast<size_t>(N) * sizeof(float);
}

extern "C" int cpfn_furthest_point_sampling(const float *xyz, int B, int N, int nsamples,
                                            int32_t *idx, void *workspace,
                                            size_t workspace_bytes, cpfn_stream_t stream) {
  using namespace cpfn;
  if (B < 0 || N < 0 || nsamples < 0) return CPFN_EINVAL;
  if (B == 0 || nsamples == 0) return CPFN_OK;
  if (N == 0 || !xyz || !idx) return CPFN_EINVAL;
  if ((static_cast<long long>(N) >> 9) >= (1 << 20)) return CPFN_EINVAL;  // rank packing
  cudaStream_t st = as_stream(stream);
  const int log2T = ref_log2_threads(N);
  const int T = 1 << log2T;
  if (N > kMaxSmemN) {
    const size_t need = cpfn_fps_workspace_bytes(B, N);
    if (!workspace || workspace_bytes < need) return CPFN_EWORKSPACE;
    fps_stream_kernel<<<B, 1024, 0, st>>>(xyz, N, nsamples, log2T,
                                          static_cast<float *>(workspace), idx);
    return check_launch();
  }
  // Block size: a multiple of T (>= one warp); N >= 1024 always uses 1024.
  const int NT = N >= 1024 ? 1024 : (T < 32 ? 32 : T);
  const int ppt = (N + NT - 1) / NT;
  if (ppt <= 1) return launch_cta<1, true>(xyz, B, N, nsamples, log2T, NT, idx, st);
  if (ppt <= 2) return launch_cta<2, true>(xyz, B, N, nsamples, log2T, NT, idx, st);
  if (ppt <= 4) return launch_cta<4, true>(xyz, B, N, nsamples, log2T, NT, idx, st);
  if (ppt <= 8) return launch_cta<8, true>(xyz, B, N, nsamples, log2T, NT, idx, st);
  return launch_cta<16, false>(xyz, B, N, nsamples, log2T, NT, idx, st);
}
