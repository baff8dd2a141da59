// Fused shared-MLP chain on the sm_100a tensor cores (tcgen05 + TMEM), with the
// neighbourhood gather / 3-NN interpolation fused in front and the max-pool fused behind.
//
// Replaces, for inference, the per-layer conv + batch-norm + ReLU (+ torch.max) sequences of
//   PointNet2/pointnet2_ops/modules/pointset_abstraction.py:61-74   (SA: group, recentre, MLP, max)
//   PointNet2/pointnet2_ops/modules/pointset_feature_propagation.py:36-51 (FP: interpolate, concat, MLP)
//   PointNet2/pn2_network.py:60-68                                  (fc1 + bn1 + ReLU + dropout + heads)
// of the reference (one cuDNN conv, one BN and one ReLU kernel per layer over 134-268 MB
// activations, plus a contiguous() copy and the 68 MB grouped tensor of SA2).
//
// One CTA owns a tile of NT columns (grouped samples, or points) and carries it through
// EVERY layer of the chain without leaving the SM:
//   worker warps  build the input tile [NT x Cin] in shared memory (gather by ball-query index
//                 and recentre / 3-point weighted interpolation + skip concat / dense rows),
//                 rounded to TF32, in the canonical K-major 128-byte-swizzled UMMA layout;
//   MMA thread    for each layer issues tcgen05.mma.kind::tf32 (M = 128 output channels, N = NT
//                 columns, K = 8 per instruction): D[ch, col] += W[ch, k] * A[col, k], with the
//                 fp32 accumulators in TMEM;
//   producer      streams the BN-folded, pre-swizzled weight blocks (16 KB = 128 ch x 32 k)
//                 through an mbarrier ring with cp.async.bulk (TMA bulk copy), running ahead
//                 across layer boundaries;
//   worker warps  read the accumulators back (tcgen05.ld, thread = channel), add the folded
//                 bias, apply ReLU (and the dropout mask), and write the activations as the NEXT
//                 layer's operand tile straight into shared memory -- or, after the last layer,
//                 max-pool over each group of columns in registers / store the rows.
// The grouped tensor, the interpolated tensor and all intermediate activations never exist in
// HBM.  Arithmetic: TF32 operands (round-to-nearest), fp32 accumulate (north_star tolerance
// for the MLP path: 1e-3 relative).
#include <string.h>

#include "common.cuh"

namespace cpfn {
namespace {

constexpr int kChainThreads = 192;   // warp 0 producer, warp 1 MMA (+ TMEM alloc), warps 2-5 workers
constexpr int kWorkers = 128;
constexpr int kStageBytes = 16384;   // one weight block: 128 rows x 128 B
constexpr int kMaxLayers = CPFN_MLP_MAX_LAYERS;
constexpr int kMaxStages = 8;

struct LayerP {
  int cin_atoms, ksteps, cout_chunks, cout, relu, bias_per_cloud, next_atoms;
  const float *bias;
  const float *mask;
  float *out_cm;
};

struct ChainP {
  int n_layers;
  LayerP L[kMaxLayers];
  const uint8_t *weights;
  int total_blocks;
  int in_mode, B, cols_per_cloud;
  long long cols;
  int n_tiles;
  const float *a_src; int a_ch, a_rows;
  const int32_t *idx;
  const float *xyz, *centers; int group_k;
  const float *b_src; int b_ch, b_rows; const float *nn_w;
  int out_mode; float *out; int ldo, pool_g;
  int act_bytes0, act_bytes1, nstage, tmem_cols;
};

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t *slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ void st_shared_f32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Shared-memory matrix descriptor, K-major, SWIZZLE_128B: start address, LBO (ignored for
// swizzled K-major) = 16 B, SBO = 1024 B (8 rows x 128 B), descriptor version 1 (Blackwell).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// Byte offset of element (row, channel) in an activation tile: K-atoms of 32 channels,
// each [NT rows x 128 B] with the 16-byte chunks XOR-swizzled by (row & 7).
template <int NT>
__device__ __forceinline__ uint32_t act_off(int row, int ch) {
  const int atom = ch >> 5, e = ch & 31;
  return static_cast<uint32_t>(atom * (NT * 128) + (row >> 3) * 1024 + (row & 7) * 128 +
                               ((((e >> 2) ^ (row & 7)) << 4) | ((e & 3) << 2)));
}

// ---- input tile builders (128 worker threads) -------------------------------------------------
template <int NT>
__device__ void load_tile(const ChainP &p, uint32_t buf, long long col0, int wq, int lane) {
  constexpr int RB = 8;                      // rows in flight per warp
  const int cin_pad = p.L[0].cin_atoms * 32;
  const int nA4 = p.a_ch >> 2;
  for (int r0 = wq * RB; r0 < NT; r0 += 4 * RB) {
    long long col[RB];
    long long arow[RB];
    bool valid[RB];
#pragma unroll
    for (int u = 0; u < RB; ++u) {
      col[u] = col0 + r0 + u;
      valid[u] = col[u] < p.cols;
      const long long c = valid[u] ? col[u] : 0;
      if (p.in_mode == CPFN_MLP_IN_GROUP) {
        const long long cloud = c / p.cols_per_cloud;
        arow[u] = cloud * p.a_rows + __ldg(p.idx + c);
      } else {
        arow[u] = c;
      }
    }
    // segment A: a_ch channels copied from a_src rows
    for (int c4 = lane; c4 < nA4; c4 += 32) {
      float4 v[RB];
#pragma unroll
      for (int u = 0; u < RB; ++u)
        v[u] = valid[u] ? __ldg(reinterpret_cast<const float4 *>(p.a_src + arow[u] * p.a_ch) + c4)
                        : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < RB; ++u) {
        const int row = r0 + u;
        const float4 t = make_float4(to_tf32(v[u].x), to_tf32(v[u].y), to_tf32(v[u].z), to_tf32(v[u].w));
        st_shared_v4(buf + act_off<NT>(row, c4 * 4), t);
      }
    }
    int done = p.a_ch;
    if (p.in_mode == CPFN_MLP_IN_GROUP) {
      // segment B: recentred position (pointset_abstraction.py:62-63): xyz[idx] - centre
      if (lane < 3) {
#pragma unroll
        for (int u = 0; u < RB; ++u) {
          float d = 0.f;
          if (valid[u]) {
            const float a = __ldg(p.xyz + arow[u] * 3 + lane);
            const float c = __ldg(p.centers + (col[u] / p.group_k) * 3 + lane);
            d = __fsub_rn(a, c);
          }
          st_shared_f32(buf + act_off<NT>(r0 + u, p.a_ch + lane), to_tf32(d));
        }
      }
      done += 3;
    } else if (p.in_mode == CPFN_MLP_IN_INTERP) {
      // segment B: three_weighted_sum (interpolate_gpu.cu:98-99): fma(p3,w3, fma(p1,w1, p2*w2))
      const int nB4 = p.b_ch >> 2;
      long long brow[RB][3];
      float w[RB][3];
#pragma unroll
      for (int u = 0; u < RB; ++u) {
        const long long c = valid[u] ? col[u] : 0;
        const long long cloud = c / p.cols_per_cloud;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          brow[u][q] = cloud * p.b_rows + __ldg(p.idx + c * 3 + q);
          w[u][q] = __ldg(p.nn_w + c * 3 + q);
        }
      }
      for (int c4 = lane; c4 < nB4; c4 += 32) {
#pragma unroll
        for (int u = 0; u < RB; ++u) {
          float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
          if (valid[u]) {
            const float4 f1 = __ldg(reinterpret_cast<const float4 *>(p.b_src + brow[u][0] * p.b_ch) + c4);
            const float4 f2 = __ldg(reinterpret_cast<const float4 *>(p.b_src + brow[u][1] * p.b_ch) + c4);
            const float4 f3 = __ldg(reinterpret_cast<const float4 *>(p.b_src + brow[u][2] * p.b_ch) + c4);
            o.x = __fmaf_rn(f3.x, w[u][2], __fmaf_rn(f1.x, w[u][0], __fmul_rn(f2.x, w[u][1])));
            o.y = __fmaf_rn(f3.y, w[u][2], __fmaf_rn(f1.y, w[u][0], __fmul_rn(f2.y, w[u][1])));
            o.z = __fmaf_rn(f3.z, w[u][2], __fmaf_rn(f1.z, w[u][0], __fmul_rn(f2.z, w[u][1])));
            o.w = __fmaf_rn(f3.w, w[u][2], __fmaf_rn(f1.w, w[u][0], __fmul_rn(f2.w, w[u][1])));
          }
          const float4 t = make_float4(to_tf32(o.x), to_tf32(o.y), to_tf32(o.z), to_tf32(o.w));
          st_shared_v4(buf + act_off<NT>(r0 + u, p.a_ch + c4 * 4), t);
        }
      }
      done += p.b_ch;
    }
    // zero the K padding (a zero weight times stale shared memory could still be NaN)
    for (int ch = done + lane; ch < cin_pad; ch += 32) {
#pragma unroll
      for (int u = 0; u < RB; ++u) st_shared_f32(buf + act_off<NT>(r0 + u, ch), 0.f);
    }
  }
}

template <int NT>
__global__ void __launch_bounds__(kChainThreads, 2)
mlp_chain_kernel(const __grid_constant__ ChainP p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t *ring = smem;
  uint8_t *act0 = ring + p.nstage * kStageBytes;
  uint8_t *act1 = act0 + p.act_bytes0;
  uint64_t *bars = reinterpret_cast<uint64_t *>(act1 + p.act_bytes1);
  uint64_t *full = bars, *empty = bars + kMaxStages, *act_ready = bars + 2 * kMaxStages,
           *acc_full = bars + 2 * kMaxStages + 1;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * kMaxStages + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.nstage; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    mbar_init(act_ready, kWorkers);
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int wave_max = p.tmem_cols / NT;

  if (warp == 0) {
    // ===== weight producer: the packed blob is consumed strictly in order, once per tile =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        for (int blk = 0; blk < p.total_blocks; ++blk) {
          mbar_wait(empty + stage, phase ^ 1);
          mbar_arrive_expect_tx(full + stage, kStageBytes);
          bulk_g2s(ring + stage * kStageBytes, p.weights + static_cast<size_t>(blk) * kStageBytes, kStageBytes,
                   full + stage);
          if (++stage == p.nstage) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(NT >> 3) << 17) |
                                 (static_cast<uint32_t>(128 >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0, act_phase = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        for (int l = 0; l < p.n_layers; ++l) {
          const LayerP &L = p.L[l];
          const uint32_t in_buf = smem_u32((l & 1) ? act1 : act0);
          for (int m0 = 0; m0 < L.cout_chunks; m0 += wave_max) {
            const int mc = min(wave_max, L.cout_chunks - m0);
            mbar_wait(act_ready, act_phase);
            act_phase ^= 1;
            tc_fence_after();
            for (int m = 0; m < mc; ++m) {
              for (int j = 0; j < L.cin_atoms; ++j) {
                mbar_wait(full + stage, phase);
                tc_fence_after();
                const uint32_t a_base = smem_u32(ring + stage * kStageBytes);
                const uint32_t b_base = in_buf + j * (NT * 128);
                const int ks = min(4, L.ksteps - 4 * j);
                for (int kk = 0; kk < ks; ++kk)
                  umma_tf32(tmem_base + m * NT, make_desc(a_base + kk * 32), make_desc(b_base + kk * 32), idesc,
                            (j | kk) != 0 ? 1u : 0u);
                umma_commit(empty + stage);          // frees the weight stage when these MMAs retire
                if (++stage == p.nstage) { stage = 0; phase ^= 1; }
              }
            }
            umma_commit(acc_full);                   // accumulators of this wave complete
          }
        }
      }
    }
  } else {
    // ===== workers: build the input tile, then the epilogue of every layer =====
    const int wq = warp & 3;                          // TMEM lane quarter this warp may access
    const int row_in_chunk = wq * 32 + lane;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      const long long col0 = static_cast<long long>(tile) * NT;
      const long long cloud = col0 / p.cols_per_cloud;
      const long long n_in_cloud = col0 - cloud * p.cols_per_cloud;
      load_tile<NT>(p, smem_u32(act0), col0, wq, lane);
      fence_proxy_async();
      mbar_arrive(act_ready);
      for (int l = 0; l < p.n_layers; ++l) {
        const LayerP &L = p.L[l];
        const bool last = (l == p.n_layers - 1);
        const uint32_t out_buf = smem_u32((l & 1) ? act0 : act1);
        const int cout_pad = L.cout_chunks * 128;
        for (int m0 = 0; m0 < L.cout_chunks; m0 += wave_max) {
          const int mc = min(wave_max, L.cout_chunks - m0);
          mbar_wait(acc_full, acc_phase);
          acc_phase ^= 1;
          tc_fence_after();
          for (int m = 0; m < mc; ++m) {
            const int ch = (m0 + m) * 128 + row_in_chunk;
            const bool ch_real = ch < L.cout;
            const float bias = __ldg(L.bias + (L.bias_per_cloud ? cloud * cout_pad : 0) + ch);
            const bool to_smem = !last && ch < L.next_atoms * 32;
            const size_t cm_base = (static_cast<size_t>(cloud) * L.cout + ch) * p.cols_per_cloud + n_in_cloud;
            float pool = 0.f;                         // post-ReLU values are >= 0
            for (int c0 = 0; c0 < NT; c0 += 32) {
              uint32_t r[32];
              tmem_ld32(tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + m * NT + c0, r);
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                float v = __uint_as_float(r[i]) + bias;
                if (L.relu) v = fmaxf(v, 0.f);
                const bool col_ok = col0 + c0 + i < p.cols;
                if (L.mask != nullptr && ch_real && col_ok) v *= __ldg(L.mask + cm_base + c0 + i);
                if (L.out_cm != nullptr && ch_real && col_ok) L.out_cm[cm_base + c0 + i] = v;
                if (to_smem) st_shared_f32(out_buf + act_off<NT>(c0 + i, ch), to_tf32(v));
                if (last) {
                  if (p.out_mode == CPFN_MLP_OUT_ROWS) {
                    if (ch_real && col_ok) p.out[(col0 + c0 + i) * p.ldo + ch] = v;
                  } else {
                    pool = fmaxf(pool, col_ok ? v : 0.f);
                    if (p.pool_g <= NT && ((c0 + i + 1) % p.pool_g) == 0) {
                      const long long grp = (col0 + c0 + i) / p.pool_g;
                      if (ch_real && col0 + c0 + i + 1 - p.pool_g < p.cols) p.out[grp * p.ldo + ch] = pool;
                      pool = 0.f;
                    }
                  }
                }
              }
            }
            if (last && p.out_mode == CPFN_MLP_OUT_POOL && p.pool_g > NT && ch_real)
              atomicMax(reinterpret_cast<int *>(p.out + (col0 / p.pool_g) * p.ldo + ch), __float_as_int(pool));
          }
          tc_fence_before();
          const bool final_wave = last && (m0 + wave_max >= L.cout_chunks);
          if (!final_wave) {
            fence_proxy_async();
            mbar_arrive(act_ready);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

int pow2_at_least(int x) {
  int p = 32;
  while (p < x) p <<= 1;
  return p;
}

template <int NT>
int launch_chain(const cpfn_mlp_chain_t *c, cudaStream_t st) {
  ChainP p{};
  p.n_layers = c->n_layers;
  int total_blocks = 0, max_chunks = 0;
  size_t act_need[2] = {0, 0};
  for (int l = 0; l < c->n_layers; ++l) {
    const cpfn_mlp_layer_t &s = c->layers[l];
    LayerP &L = p.L[l];
    if (s.cin <= 0 || s.cout <= 0 || !s.bias) return CPFN_EINVAL;
    L.cin_atoms = (s.cin + 31) / 32;
    L.ksteps = (s.cin + 7) / 8;
    L.cout_chunks = (s.cout + 127) / 128;
    L.cout = s.cout;
    L.relu = s.relu;
    L.bias_per_cloud = s.bias_per_cloud;
    L.bias = s.bias; L.mask = s.mask; L.out_cm = s.out_cm;
    L.next_atoms = 0;
    if (l > 0) {
      if (s.cin != c->layers[l - 1].cout) return CPFN_EINVAL;
      p.L[l - 1].next_atoms = L.cin_atoms;
    }
    total_blocks += L.cout_chunks * L.cin_atoms;
    if (L.cout_chunks > max_chunks) max_chunks = L.cout_chunks;
    const size_t in_bytes = static_cast<size_t>(L.cin_atoms) * NT * 128;
    if (in_bytes > act_need[l & 1]) act_need[l & 1] = in_bytes;
    if ((s.bias_per_cloud || s.mask || s.out_cm) && (c->cols_per_cloud % NT) != 0) return CPFN_EINVAL;
  }
  if (static_cast<size_t>(total_blocks) * kStageBytes != c->weight_bytes) return CPFN_EINVAL;
  p.weights = static_cast<const uint8_t *>(c->weights);
  p.total_blocks = total_blocks;
  p.in_mode = c->in_mode; p.B = c->B; p.cols_per_cloud = c->cols_per_cloud;
  p.cols = static_cast<long long>(c->B) * c->cols_per_cloud;
  p.n_tiles = static_cast<int>((p.cols + NT - 1) / NT);
  p.a_src = c->a_src; p.a_ch = c->a_ch; p.a_rows = c->a_rows;
  p.idx = c->idx; p.xyz = c->xyz; p.centers = c->centers; p.group_k = c->group_k;
  p.b_src = c->b_src; p.b_ch = c->b_ch; p.b_rows = c->b_rows; p.nn_w = c->nn_w;
  p.out_mode = c->out_mode; p.out = c->out; p.ldo = c->ldo; p.pool_g = c->pool_g;
  // input row width must match layer 0
  int width = c->a_ch;
  if (c->in_mode == CPFN_MLP_IN_GROUP) width += 3;
  else if (c->in_mode == CPFN_MLP_IN_INTERP) width += c->b_ch;
  if (width != c->layers[0].cin || (c->a_ch & 3) || (c->b_ch & 3)) return CPFN_EINVAL;
  if (c->a_ch > 0 && !c->a_src) return CPFN_EINVAL;
  if (c->in_mode == CPFN_MLP_IN_GROUP && (!c->idx || !c->xyz || !c->centers || c->group_k <= 0)) return CPFN_EINVAL;
  if (c->in_mode == CPFN_MLP_IN_INTERP && (!c->idx || !c->b_src || !c->nn_w)) return CPFN_EINVAL;
  if (c->out_mode == CPFN_MLP_OUT_POOL) {
    if (c->pool_g <= 0 || (c->pool_g % 32) != 0 || !c->layers[c->n_layers - 1].relu) return CPFN_EINVAL;
    if (c->pool_g <= NT ? (NT % c->pool_g) != 0 : ((c->pool_g % NT) != 0 || (c->cols_per_cloud % c->pool_g) != 0))
      return CPFN_EINVAL;
  }
  p.tmem_cols = pow2_at_least(NT * (max_chunks < 512 / NT ? max_chunks : 512 / NT));
  p.act_bytes0 = static_cast<int>(act_need[0]);
  p.act_bytes1 = static_cast<int>(act_need[1]);
  const size_t fixed = act_need[0] + act_need[1] + 1024 /*align*/ + 256 /*barriers*/;
  const size_t max_smem = 227 * 1024;
  if (fixed + 2 * kStageBytes > max_smem) return CPFN_EINVAL;
  // Two CTAs per SM (one's epilogue overlaps the other's MMAs) when shared memory and TMEM allow.
  int per_sm = 1;
  int nstage = static_cast<int>((max_smem - fixed) / kStageBytes);
  if (p.tmem_cols <= 256 && fixed + 2 * kStageBytes <= max_smem / 2 - 1024) {
    per_sm = 2;
    nstage = static_cast<int>((max_smem / 2 - 1024 - fixed) / kStageBytes);
  }
  if (nstage > kMaxStages) nstage = kMaxStages;
  if (nstage > total_blocks) nstage = total_blocks < 2 ? 2 : total_blocks;
  p.nstage = nstage;
  const size_t smem = fixed + static_cast<size_t>(nstage) * kStageBytes;
  auto kern = mlp_chain_kernel<NT>;
  CPFN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const int sms = sm_count() > 0 ? sm_count() : 148;
  const int grid = p.n_tiles < per_sm * sms ? p.n_tiles : per_sm * sms;
  if (grid <= 0) return CPFN_OK;
  if (c->out_mode == CPFN_MLP_OUT_POOL && c->pool_g > NT)
    CPFN_CUDA_TRY(cudaMemsetAsync(c->out, 0, sizeof(float) * static_cast<size_t>(p.cols / c->pool_g) * c->ldo, st));
  kern<<<grid, kChainThreads, smem, st>>>(p);
  return check_launch();
}

}  // namespace
}  // namespace cpfn

// Host-side layout transform: W [cout, cin] row-major fp32 (BatchNorm already folded) ->
// blocks of 128 output channels x 32 input channels in the kernel's shared-memory image
// (TF32 round-to-nearest-away, zero padded, 16-byte chunks XOR-swizzled by row & 7),
// ordered chunk-major then K-atom -- exactly the order the MMA thread consumes them.
extern "C" size_t cpfn_mlp_packed_bytes(int cout, int cin) {
  if (cout <= 0 || cin <= 0) return 0;
  return static_cast<size_t>((cout + 127) / 128) * ((cin + 31) / 32) * cpfn::kStageBytes;
}

extern "C" int cpfn_mlp_pack_weights_host(const float *W, int cout, int cin, void *packed) {
  if (!W || !packed || cout <= 0 || cin <= 0) return CPFN_EINVAL;
  const int chunks = (cout + 127) / 128, atoms = (cin + 31) / 32;
  uint32_t *dst = static_cast<uint32_t *>(packed);
  for (int m = 0; m < chunks; ++m)
    for (int j = 0; j < atoms; ++j) {
      uint32_t *blk = dst + (static_cast<size_t>(m) * atoms + j) * (cpfn::kStageBytes / 4);
      for (int r = 0; r < 128; ++r)
        for (int e = 0; e < 32; ++e) {
          const int co = m * 128 + r, ci = j * 32 + e;
          uint32_t bits = 0;
          if (co < cout && ci < cin) {
            float f = W[static_cast<size_t>(co) * cin + ci];
            uint32_t u;
            memcpy(&u, &f, 4);
            if ((u & 0x7F800000u) != 0x7F800000u) u += 0x1000u;   // cvt.rna.tf32: nearest, ties away
            bits = u & 0xFFFFE000u;
          }
          const int off = (r >> 3) * 1024 + (r & 7) * 128 + ((((e >> 2) ^ (r & 7)) << 4) | ((e & 3) << 2));
          blk[off >> 2] = bits;
        }
    }
  return CPFN_OK;
}

extern "C" int cpfn_mlp_chain(const cpfn_mlp_chain_t *c, cpfn_stream_t stream) {
  using namespace cpfn;
  if (!c || c->n_layers <= 0 || c->n_layers > kMaxLayers || c->B < 0 || c->cols_per_cloud < 0) return CPFN_EINVAL;
  if (c->B == 0 || c->cols_per_cloud == 0) return CPFN_OK;
  if (!c->weights || !c->out) return CPFN_EINVAL;
  cudaStream_t st = as_stream(stream);
  switch (c->tile_cols) {
    case 128: return launch_chain<128>(c, st);
    case 64: return launch_chain<64>(c, st);
    case 32: return launch_chain<32>(c, st);
    default: return CPFN_EINVAL;
  }
}
