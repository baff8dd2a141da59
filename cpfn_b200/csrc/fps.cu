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
#include <stdlib.h>

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

// Stage xyz [N,3] (AoS, global) into shared SoA sx|sy|sz with pitch Np.  16-byte loads, eight in flight per thread
// (a cluster CTA stages its whole 96 KB cloud before the first round; with scalar loads that was ~7 us of every
// launch).
__device__ __forceinline__ void stage_soa(const float *__restrict__ p, int N, int Np,
                                          float *s) {
  const int total = 3 * N;
  if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
    const int nv = total >> 2;
    const float4 *pv = reinterpret_cast<const float4 *>(p);
    for (int v0 = threadIdx.x; v0 < nv; v0 += 8 * blockDim.x) {
      float4 q[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int v = v0 + u * blockDim.x;
        q[u] = v < nv ? __ldg(pv + v) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int v = v0 + u * blockDim.x;
        if (v < nv) {
          const int f = 4 * v;
          const float e[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            const int k = (f + w) / 3, c = (f + w) - 3 * k;
            s[c * Np + k] = e[w];
          }
        }
      }
    }
    for (int f = (nv << 2) + threadIdx.x; f < total; f += blockDim.x) {
      const int k = f / 3, c = f - 3 * k;
      s[c * Np + k] = __ldg(p + f);
    }
    return;
  }
  for (int f = threadIdx.x; f < total; f += blockDim.x) {
    const int k = f / 3, c = f - 3 * k;
    s[c * Np + k] = __ldg(p + f);
  }
}

// PPT points per thread, thread t owns k = t + i*blockDim.x.  blockDim.x is a
// multiple of T (or PPT == 1), so scanning i upward visits a thread's points
// in increasing reference rank and the strict '>' keeps the right one on ties.
template <int PPT, bool COORDS_IN_REGS>
__global__ void __launch_bounds__(1024, 1)
fps_cta_kernel(const float *__restrict__ xyz, float *__restrict__ new_xyz, int N, int m, int log2T,
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
  float *nx = new_xyz ? new_xyz + static_cast<size_t>(blockIdx.x) * m * 3 : nullptr;   // the centroids themselves
  if (nx && t == 0) { nx[0] = cx; nx[1] = cy; nx[2] = cz; }
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
    if (nx && t == 0) { nx[3 * j] = cx; nx[3 * j + 1] = cy; nx[3 * j + 2] = cz; }
  }
}

// ---- thread-block-cluster variant ---------------------------------------------------------------
// At GlobalSPFN's B = 16 clouds the one-CTA kernel above uses 16 of 148 SMs and every round costs
// the whole cloud's 8192 distance updates on ONE SM.  Here a cluster of C CTAs (C = 8, 4 or 2,
// chosen so that B*C <= #SMs) owns a cloud: CTA r keeps points [r*Nc, (r+1)*Nc) and their running
// minima in registers, reduces them to one 64-bit (distance, ~rank) key, and pushes that key into the
// shared memory of all C CTAs with st.async (DSMEM), whose completion is counted by an mbarrier in
// the RECEIVING CTA -- no cluster-wide barrier (~380 cycles) in the round loop.  Every CTA then
// takes the maximum of the C keys, which is the same winner (and the same tie-break) as the
// reference's single block.  Keys and barriers are double-buffered by round parity: a CTA can only
// be one round ahead of its slowest peer, because round j+1 needs every peer's round-j key.
__device__ __forceinline__ uint32_t fps_smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_async_u64(uint32_t remote_addr, unsigned long long v, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];"
               ::"r"(remote_addr), "l"(v), "r"(remote_bar) : "memory");
}
__device__ __forceinline__ void fps_mbar_init(unsigned long long *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(fps_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fps_mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fps_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fps_mbar_wait_cluster(unsigned long long *bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "FPS_WAIT:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra FPS_DONE;\n\t"
      "bra FPS_WAIT;\n\t"
      "FPS_DONE:\n\t"
      "}\n" ::"r"(fps_smem_u32(bar)), "r"(parity) : "memory");
}

// Polling exchange (no mbarrier): a key carries a 2-bit round tag in bits 30-31 of its low word (the rank needs 29
// bits and its complement always has bit 29 set), the sender stores it into the peer's shared memory with a plain
// DSMEM store and the receiver spins on its OWN shared memory until every slot shows the tag of the round.  The
// mbarrier form costs one complete_tx transaction per 8-byte message at the receiver (32-64 per round), which is
// what bounded the round time.
__device__ __forceinline__ void st_cluster_u64(uint32_t remote_addr, unsigned long long v) {
  asm volatile("st.relaxed.cluster.shared::cluster.b64 [%0], %1;" ::"r"(remote_addr), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_volatile_shared_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.volatile.shared::cta.b64 %0, [%1];" : "=l"(v) : "r"(fps_smem_u32(p)) : "memory");
  return v;
}
// Key low word of the cluster kernels (N <= 16384, reference block T = 512): bits 14-27 = the complement of the
// 14-bit reference rank (bitrev9(k mod 512) << 5 | k >> 9: larger = wins ties), bits 0-13 = k itself, so the
// arg-max over (distance, code) is the reference's winner and its index needs no decoding.  Never 0 for a point.
__device__ __forceinline__ uint32_t fps_code(uint32_t k) {
  const uint32_t rank14 = ((__brev(k & 511u) >> 23) << 5) | (k >> 9);
  return ((~rank14 & 0x3FFFu) << 14) | (k & 0x3FFFu);
}
__device__ __forceinline__ uint32_t fps_tag(int j) { return ((static_cast<uint32_t>(j - 1) >> 1) & 1u) + 1u; }

// Waits until this lane's (up to two) slots carry `tag`, returns the larger key with the tag stripped.
__device__ __forceinline__ unsigned long long fps_poll_keys(const unsigned long long *slots, int lane, int nkeys,
                                                            uint32_t tag) {
  unsigned long long v0 = 0ull, v1 = 0ull;
  while (true) {
    bool ok = true;
    if (lane < nkeys) { v0 = ld_volatile_shared_u64(slots + lane); ok = ((static_cast<uint32_t>(v0) >> 30) == tag); }
    if (lane + 32 < nkeys) { v1 = ld_volatile_shared_u64(slots + lane + 32); ok = ok && ((static_cast<uint32_t>(v1) >> 30) == tag); }
    if (__all_sync(0xffffffffu, ok)) break;
  }
  const unsigned long long v = v0 > v1 ? v0 : v1;
  return v & 0xFFFFFFFF3FFFFFFFull;
}

// Round-phase profile of fps_cluster_kernel (cycles summed over the rounds by thread 0 of CTA 0 when the kernel is
// built into its PROFILE form): read back with cpfn_debug_fps_profile.
__device__ long long g_fps_prof[8];

constexpr int kFpsClusterThreads = 256;
constexpr int kFpsClusterWarps = kFpsClusterThreads / 32;

// Thread t owns k = r*Nc + t + i*256 (Nc a multiple of 512).  With the reference's T = 512 the rank of
// point i is (bitrev9(k mod 512) << 20) | (k >> 9): even i share k mod 512 = t, odd i have t + 256 whose
// bit-reversal is one larger, so scanning the even i first and then the odd i visits a thread's points
// in increasing reference rank and the strict '>' keeps the right one on ties.
template <int PPT, bool POLL, bool PROFILE = false>
__global__ void __launch_bounds__(kFpsClusterThreads, 1)
fps_cluster_kernel(const float *__restrict__ xyz, float *__restrict__ new_xyz, int N, int m, int log2T, int Nc,
                   int32_t *__restrict__ idx, int j_begin, int j_end, float *__restrict__ state) {
  extern __shared__ float s_xyz[];
  __shared__ __align__(8) unsigned long long xslot[2][8 * kFpsClusterWarps];   // [parity][sender CTA * 8 + sender warp]
  __shared__ __align__(8) unsigned long long xbar[2];
  const uint32_t C = cluster_nctarank(), r = cluster_ctarank();
  const int cloud = blockIdx.x / C;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  constexpr int NT = kFpsClusterThreads;
  const int Np = (N + 31) & ~31;
  const float *p = xyz + static_cast<size_t>(cloud) * N * 3;
  int32_t *out = idx + static_cast<size_t>(cloud) * m;
  float *sx = s_xyz, *sy = s_xyz + Np, *sz = s_xyz + 2 * Np;

  stage_soa(p, N, Np, s_xyz);            // every CTA keeps the whole cloud for the centroid look-up
  if (t == 0) {
    fps_mbar_init(&xbar[0], 1);
    fps_mbar_init(&xbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = t; i < 2 * 8 * kFpsClusterWarps; i += NT) (&xslot[0][0])[i] = 0ull;
  __syncthreads();
  cluster_sync_all();                    // all barriers exist before any peer can complete_tx on them

  float px[PPT], py[PPT], pz[PPT], tmp[PPT];
  uint32_t code[PPT];
  const int k0 = static_cast<int>(r) * Nc + t;
#pragma unroll
  for (int i = 0; i < PPT; ++i) {
    const int k = k0 + i * NT;
    float x = 0.f, y = 0.f, z = 0.f;
    bool live = false;
    if (k < N && i * NT + t < Nc) {
      x = sx[k]; y = sy[k]; z = sz[k];
      live = !(static_cast<double>(sqnorm3(x, y, z)) <= 1e-3);
    }
    px[i] = x; py[i] = y; pz[i] = z;
    tmp[i] = live ? 1e10f : -1.0f;
    // resumed launch (rounds j_begin .. j_end-1 of a sampling split over several launches): the running minima of
    // the rounds before come back from `state`
    if (j_begin > 1 && live) tmp[i] = state[static_cast<size_t>(cloud) * N + k];
    code[i] = fps_code(static_cast<uint32_t>(k));      // tie-break order and the index itself, computed once
  }
  // lane l < C of every warp delivers the warp's key to CTA l: slot [parity][r*8 + warp], barrier [parity]
  uint32_t my_slot0 = 0, my_slot1 = 0, my_bar0 = 0, my_bar1 = 0;
  if (lane < static_cast<int>(C)) {
    my_slot0 = mapa_u32(fps_smem_u32(&xslot[0][r * kFpsClusterWarps + warp]), lane);
    my_slot1 = mapa_u32(fps_smem_u32(&xslot[1][r * kFpsClusterWarps + warp]), lane);
    my_bar0 = mapa_u32(fps_smem_u32(&xbar[0]), lane);
    my_bar1 = mapa_u32(fps_smem_u32(&xbar[1]), lane);
  }
  const int nkeys = static_cast<int>(C) * kFpsClusterWarps;      // <= 64: two per lane

  float *nx = (new_xyz && r == 0 && t == 0) ? new_xyz + static_cast<size_t>(cloud) * m * 3 : nullptr;
  float cx, cy, cz;
  if (j_begin <= 1) {
    if (r == 0 && t == 0) out[0] = 0;
    cx = sx[0]; cy = sy[0]; cz = sz[0];
    if (nx) { nx[0] = cx; nx[1] = cy; nx[2] = cz; }
  } else {
    const int last = __ldg(out + j_begin - 1);             // written by the launch before this one
    cx = sx[last]; cy = sy[last]; cz = sz[last];
  }
  long long pt[6] = {0, 0, 0, 0, 0, 0}, c0 = 0;
#define FPS_MARK(i) if (PROFILE) { const long long c1 = clock64(); pt[i] += c1 - c0; c0 = c1; }
  if (PROFILE) c0 = clock64();
  for (int j = max(1, j_begin); j < j_end; ++j) {
    // running minima, then the thread's arg-max in the reference's order (even i first, then odd i: increasing
    // reference rank; of equal distances the earlier one wins).  "Later wins only if strictly greater" is associative,
    // so the scan is a balanced tree over that order: 3 dependent compare / select levels for 8 points instead of 8.
    float dv[PPT];
    uint32_t cv[PPT];
#pragma unroll
    for (int q = 0; q < PPT; ++q) {
      const int i = (q < (PPT + 1) / 2) ? 2 * q : 2 * (q - (PPT + 1) / 2) + 1;
      const float d2 = fminf(sqdist3(px[i], py[i], pz[i], cx, cy, cz), tmp[i]);
      tmp[i] = d2;
      dv[q] = d2;
      cv[q] = code[i];
    }
#pragma unroll
    for (int w = 1; w < PPT; w <<= 1) {
#pragma unroll
      for (int q = 0; q + w < PPT; q += 2 * w) {
        const bool later = dv[q + w] > dv[q];
        dv[q] = later ? dv[q + w] : dv[q];
        cv[q] = later ? cv[q + w] : cv[q];
      }
    }
    const float best = dv[0] > -1.0f ? dv[0] : -1.0f;      // all points dead (-1): no candidate, as the scan from -1
    const uint32_t bc = dv[0] > -1.0f ? cv[0] : 0u;
    uint32_t hi = 0u, lo = 0u;
    if (best >= 0.0f) {
      hi = __float_as_uint(best);
      lo = bc;
    }
    FPS_MARK(0)                                            // distance update
    const uint32_t Mw = __reduce_max_sync(0xffffffffu, hi);
    const uint32_t Lw = __reduce_max_sync(0xffffffffu, hi == Mw ? lo : 0u);
    const int par = j & 1;
    unsigned long long v;
    FPS_MARK(1)                                            // warp arg-max (2 REDUX)
    if (POLL) {
      const uint32_t tag = fps_tag(j);
      const unsigned long long key = (static_cast<unsigned long long>(Mw) << 32) | Lw | (tag << 30);
      if (lane < static_cast<int>(C)) st_cluster_u64(par ? my_slot1 : my_slot0, key);
      FPS_MARK(2)                                          // push
      v = fps_poll_keys(&xslot[par][0], lane, nkeys, tag);
      FPS_MARK(3)                                          // wait for all keys
    } else {
      const unsigned long long key = (static_cast<unsigned long long>(Mw) << 32) | Lw;
      if (t == 0) fps_mbar_expect_tx(&xbar[par], 8u * nkeys);
      if (lane < static_cast<int>(C)) st_async_u64(par ? my_slot1 : my_slot0, key, par ? my_bar1 : my_bar0);
      fps_mbar_wait_cluster(&xbar[par], ((j - 1) >> 1) & 1);   // (j-1)/2-th use of this barrier
      const unsigned long long v0 = lane < nkeys ? xslot[par][lane] : 0ull;
      const unsigned long long v1 = lane + 32 < nkeys ? xslot[par][lane + 32] : 0ull;
      v = v0 > v1 ? v0 : v1;
    }
    const uint32_t h2 = static_cast<uint32_t>(v >> 32), l2 = static_cast<uint32_t>(v);
    const uint32_t M = __reduce_max_sync(0xffffffffu, h2);
    const uint32_t L = __reduce_max_sync(0xffffffffu, h2 == M ? l2 : 0u);
    const int old = static_cast<int>(L & 0x3FFFu);          // the low 14 bits of a code are the point index (0 if none)
    FPS_MARK(4)                                            // cluster arg-max (2 REDUX) + unrank
    if (r == 0 && t == 0) out[j] = old;
    cx = sx[old]; cy = sy[old]; cz = sz[old];
    if (nx) { nx[3 * j] = cx; nx[3 * j + 1] = cy; nx[3 * j + 2] = cz; }
    if (PROFILE) { if (cx + cy + cz == 12345.f) pt[5] += 1; }
    FPS_MARK(5)                                            // centroid look-up
  }
#undef FPS_MARK
  if (state != nullptr && j_end < m) {                     // more rounds follow in another launch
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
      const int k = k0 + i * NT;
      if (k < N && i * NT + t < Nc) state[static_cast<size_t>(cloud) * N + k] = tmp[i];
    }
  }
  if (PROFILE && blockIdx.x == 0 && t == 0)
    for (int i = 0; i < 6; ++i) g_fps_prof[i] = pt[i];
  cluster_sync_all();                    // nobody leaves while a peer may still write into its shared memory
}

template <int PPT>
int launch_cluster(const float *xyz, float *new_xyz, int B, int N, int m, int log2T, int C, int Nc, int32_t *idx,
                   cudaStream_t st, int j_begin = 0, int j_end = -1, float *state = nullptr, size_t smem_floor = 0) {
  size_t smem = 3u * static_cast<size_t>((N + 31) & ~31) * sizeof(float);
  if (smem < smem_floor) smem = smem_floor;      // keeps other kernels' CTAs off the sampling SMs (see ..._rounds)
  if (j_end < 0) j_end = m;
  const char *poll_env = getenv("CPFN_FPS_POLL");
  auto kern = (poll_env && poll_env[0] == '0') ? fps_cluster_kernel<PPT, false> : fps_cluster_kernel<PPT, true>;
  if (getenv("CPFN_FPS_PROFILE") != nullptr) kern = fps_cluster_kernel<PPT, true, true>;
  CPFN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(B) * C);
  cfg.blockDim = dim3(kFpsClusterThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CPFN_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, xyz, new_xyz, N, m, log2T, Nc, idx, j_begin, j_end, state));
  return check_launch();
}

// Any N: running minimum in a global workspace [B,N], coordinates re-read
// through L1/L2 every round.  Only used beyond kMaxSmemN points per cloud.
__global__ void __launch_bounds__(1024, 1)
fps_stream_kernel(const float *__restrict__ xyz, float *__restrict__ new_xyz, int N, int m, int log2T,
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
  float *nx = (new_xyz && t == 0) ? new_xyz + static_cast<size_t>(blockIdx.x) * m * 3 : nullptr;
  if (nx) { nx[0] = cx; nx[1] = cy; nx[2] = cz; }
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
    if (nx) { nx[3 * j] = cx; nx[3 * j + 1] = cy; nx[3 * j + 2] = cz; }
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
int launch_cta(const float *xyz, float *new_xyz, int B, int N, int m, int log2T, int NT, int32_t *idx,
               cudaStream_t st) {
  const size_t smem = 3u * static_cast<size_t>((N + 31) & ~31) * sizeof(float);
  auto kern = fps_cta_kernel<PPT, REGS>;
  if (smem + 1024 > 48 * 1024)  // static shared memory counts against the 48 KB default
    CPFN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem)));
  kern<<<B, NT, smem, st>>>(xyz, new_xyz, N, m, log2T, idx);
  return check_launch();
}

}  // namespace
}  // namespace cpfn

extern "C" int cpfn_debug_fps_profile(long long *cycles6) {
  if (!cycles6) return CPFN_EINVAL;
  CPFN_CUDA_TRY(cudaMemcpyFromSymbol(cycles6, cpfn::g_fps_prof, 6 * sizeof(long long)));
  return CPFN_OK;
}

extern "C" size_t cpfn_fps_workspace_bytes(int B, int N) {
  if (B <= 0 || N <= cpfn::kMaxSmemN) return 0;
  return static_cast<size_t>(B) * static_cast<size_t>(N) * sizeof(float);
}

extern "C" int cpfn_fps_rounds_supported(int B, int N) {
  using namespace cpfn;
  if (B <= 0 || N < 2048 || N > kMaxSmemN || (1 << ref_log2_threads(N)) != 512) return 0;
  const int sms = sm_count() > 0 ? sm_count() : 148;
  int C = 8;
  while (C > 1 && N / C < 2048) C >>= 1;
  while (C > 1 && static_cast<long long>(B) * C > sms) C >>= 1;
  return C > 1 ? 1 : 0;
}

extern "C" int cpfn_furthest_point_sampling_rounds(const float *xyz, int B, int N, int nsamples, int j_begin, int j_end,
                                                   int32_t *idx, float *new_xyz, float *state, size_t state_bytes,
                                                   size_t smem_floor_bytes, cpfn_stream_t stream) {
  using namespace cpfn;
  if (B <= 0 || nsamples <= 0 || j_begin < 0 || j_end > nsamples || j_begin >= j_end || !xyz || !idx) return CPFN_EINVAL;
  if (!cpfn_fps_rounds_supported(B, N)) return CPFN_EINVAL;
  if ((j_end < nsamples || j_begin > 0) && (!state || state_bytes < sizeof(float) * static_cast<size_t>(B) * N))
    return CPFN_EWORKSPACE;
  if (smem_floor_bytes > 200u * 1024u) return CPFN_EINVAL;
  const int sms = sm_count() > 0 ? sm_count() : 148;
  int C = 8;
  while (C > 1 && N / C < 2048) C >>= 1;
  while (C > 1 && static_cast<long long>(B) * C > sms) C >>= 1;
  const int log2T = ref_log2_threads(N);
  const int Nc = ((N + C - 1) / C + 511) / 512 * 512;
  const int ppt = Nc / 256;
  cudaStream_t st = as_stream(stream);
  if (ppt <= 2) return launch_cluster<2>(xyz, new_xyz, B, N, nsamples, log2T, C, Nc, idx, st, j_begin, j_end, state, smem_floor_bytes);
  if (ppt <= 4) return launch_cluster<4>(xyz, new_xyz, B, N, nsamples, log2T, C, Nc, idx, st, j_begin, j_end, state, smem_floor_bytes);
  if (ppt <= 8) return launch_cluster<8>(xyz, new_xyz, B, N, nsamples, log2T, C, Nc, idx, st, j_begin, j_end, state, smem_floor_bytes);
  if (ppt <= 16) return launch_cluster<16>(xyz, new_xyz, B, N, nsamples, log2T, C, Nc, idx, st, j_begin, j_end, state, smem_floor_bytes);
  return CPFN_EINVAL;
}

extern "C" int cpfn_furthest_point_sampling(const float *xyz, int B, int N, int nsamples,
                                            int32_t *idx, void *workspace,
                                            size_t workspace_bytes, cpfn_stream_t stream) {
  return cpfn_furthest_point_sampling_xyz(xyz, B, N, nsamples, idx, nullptr, workspace, workspace_bytes, stream);
}

extern "C" int cpfn_furthest_point_sampling_xyz(const float *xyz, int B, int N, int nsamples, int32_t *idx,
                                                float *new_xyz, void *workspace, size_t workspace_bytes,
                                                cpfn_stream_t stream) {
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
    fps_stream_kernel<<<B, 1024, 0, st>>>(xyz, new_xyz, N, nsamples, log2T,
                                          static_cast<float *>(workspace), idx);
    return check_launch();
  }
  // Few clouds, many points: a cluster of C CTAs per cloud (B*C <= #SMs), 512 threads each.
  if (N >= 2048 && T == 512 && getenv("CPFN_FPS_NO_CLUSTER") == nullptr) {
    const int sms = sm_count() > 0 ? sm_count() : 148;
    // about 2048 points per CTA (8 per thread) balances the per-round compute against the key exchange,
    // whose cost grows with the cluster size (measured at B=16, N=8192: C=8 325 us, C=4 268 us, C=2 297 us)
    int C = 8;
    while (C > 1 && N / C < 2048) C >>= 1;
    if (const char *e = getenv("CPFN_FPS_CLUSTER")) C = atoi(e) > 0 ? atoi(e) : C;   // tuning knob
    while (C > 1 && static_cast<long long>(B) * C > sms) C >>= 1;
    if (C > 1) {
      const int Nc = ((N + C - 1) / C + 511) / 512 * 512;
      const int ppt = Nc / 256;
      if (ppt <= 2) return launch_cluster<2>(xyz, new_xyz, B, N, nsamples, log2T, C, Nc, idx, st);
      if (ppt <= 4) return launch_cluster<4>(xyz, new_xyz, B, N, nsamples, log2T, C, Nc, idx, st);
      if (ppt <= 8) return launch_cluster<8>(xyz, new_xyz, B, N, nsamples, log2T, C, Nc, idx, st);
      if (ppt <= 16) return launch_cluster<16>(xyz, new_xyz, B, N, nsamples, log2T, C, Nc, idx, st);
    }
  }
  // Block size: a multiple of T (>= one warp); N >= 1024 always uses 1024.
  const int NT = N >= 1024 ? 1024 : (T < 32 ? 32 : T);
  const int ppt = (N + NT - 1) / NT;
  if (ppt <= 1) return launch_cta<1, true>(xyz, new_xyz, B, N, nsamples, log2T, NT, idx, st);
  if (ppt <= 2) return launch_cta<2, true>(xyz, new_xyz, B, N, nsamples, log2T, NT, idx, st);
  if (ppt <= 4) return launch_cta<4, true>(xyz, new_xyz, B, N, nsamples, log2T, NT, idx, st);
  if (ppt <= 8) return launch_cta<8, true>(xyz, new_xyz, B, N, nsamples, log2T, NT, idx, st);
  return launch_cta<16, false>(xyz, new_xyz, B, N, nsamples, log2T, NT, idx, st);
}
