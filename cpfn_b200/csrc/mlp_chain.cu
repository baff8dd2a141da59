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
//                 and recentre / 3-point weighted interpolation + skip concat / dense rows) in the
//                 canonical K-major 128-byte-swizzled UMMA layout;
//   MMA warp      for each layer issues tcgen05.mma.kind::f16 (K = 16 per instruction) with the fp32
//                 accumulators in TMEM.  The warp runs its loops warp-uniformly and predicates only the
//                 tcgen05 instructions on elect.sync, so descriptors and addresses live in uniform
//                 registers (inside an `if (lane == 0)` ptxas wrapped every UTCHMMA in an ELECT /
//                 R2UR.BROADCAST / BRA.U.ANY loop: ~200 cycles per MMA issued);
//   producer      streams the BN-folded, pre-swizzled weight blocks (16 KB = 128 ch x 64 k bf16)
//                 through an mbarrier ring with cp.async.bulk (TMA bulk copy), running ahead
//                 across layer boundaries;
//   worker warps  read the accumulators back (tcgen05.ld), add the folded bias, apply ReLU (and the
//                 dropout bits), and write the activations as the NEXT layer's operand tile straight
//                 into shared memory -- or, after the last layer, max-pool over each group of columns
//                 in registers / store the rows.
// Two operand orientations: channels = M (mlp_chain_kernel<64|32>: thread = channel in the epilogue) and
// points = M (mlp_chain_pm_kernel / mlp_chain_pm1_kernel: thread = point, half the epilogue instructions per
// element); the points-as-M kernels switch to channels = M for a POOLED last layer (weights as the A
// operand), so that the pooling is an in-register max over TMEM columns (epi_pool_cols).
// The grouped tensor, the interpolated tensor and all intermediate activations never exist in
// HBM.  CPFN_CHAIN_PROFILE=1 makes every launch record where its MMA warp, its producer and a worker
// warp spend their cycles (Prof, cpfn_debug_chain_profile).
//
// Arithmetic: split-bf16 ("bf16x3").  Every fp32 operand x is held as hi = bf16(x) and
// lo = bf16(x - hi) (together 16 mantissa bits) and each product is accumulated in fp32 as
// a_hi*w_hi + a_lo*w_hi + a_hi*w_lo: three bf16 MMAs at twice the TF32 rate, i.e. 3/4 of the
// tensor time of one TF32 pass and the same shared-memory footprint as fp32 operands, with a
// per-product error of ~2^-16 instead of TF32's 2^-11.  Measured end-to-end error of the
// 14-layer GlobalSPFN forward against the fp32 oracle: ~1e-5 of the tensor scale (plain TF32:
// 2e-3 .. 5e-3, which misses the 1e-3 tolerance of the north star).
#include <stdlib.h>
#include <string.h>

#include <cuda_bf16.h>

#include "common.cuh"

namespace cpfn {
namespace {

constexpr int kChainThreads = 320;   // warp 0 producer, warp 1 MMA (+ TMEM alloc), warps 2-9 workers
constexpr int kWorkers = 256;
constexpr int kStageBytes = 16384;   // one weight block: 128 rows x 128 B (64 bf16)
constexpr int kMaxLayers = CPFN_MLP_MAX_LAYERS;
constexpr int kMaxStages = 8;
constexpr int kMiscBytes = 10240;    // barriers, TMEM slot, per-row loader scratch (2 x 1152 words in the two-sub-tile kernel)

struct LayerP {
  int cin_atoms, ksteps, cout_chunks, cout, relu, bias_per_cloud, next_k16;
  const float *bias;
  const uint32_t *mask_bits;   // dropout keep bits, [col][mask_words] (bit ch%32 of word ch/32)
  float mask_scale;
  int mask_words;
  float *out_cm;
};

struct ChainP {
  int split_cout;     // single-layer chains only: gridDim.y = cout chunks, one chunk per CTA
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
  const float *l0_w, *l0_b; int l0_cout;   // GROUP mode without features: first layer (3 -> l0_cout) on CUDA cores
  const float *in_bias;                    // INTERP mode: relu(interpolated + in_bias) enters layer 0
  int win_cols, win_off;                   // column window: win_cols columns of every cloud, from column win_off (0: all)
  int pool_fast;                           // pooled last layer with swapped operand roles (see epi_pool_cols)
  const float *xyz_w;                      // GROUP + features: layer 0's position columns as an epilogue term [cout_pad][4]
  int act_bytes0, act_bytes1, nstage, tmem_cols;
  unsigned long long *prof;                // phase profile of this launch (kProfSlots sums), or nullptr
  uint32_t wait_hint;                      // suspend-time hint (ns) of the mbarrier waits (CPFN_CHAIN_WAIT_HINT)
};

// Phase profile (CPFN_CHAIN_PROFILE=1, read back with cpfn_debug_chain_profile): clock64 sums over the CTAs of a launch.
//   0 MMA thread waiting for the operand tile   1 MMA thread waiting for weight blocks   2 MMA thread, whole loop
//   3 weight producer waiting for a free stage  4 weight producer, whole loop
//   5 first worker warp building the input tile 6 ... waiting for the accumulators       7 ... whole loop
//   8 CTAs                                      9 tiles (sub-tiles) processed by the profiled worker warps
//   10 + l: the profiled worker warp's epilogue of layer l (l <= 5)
constexpr int kProfSlots = 16;
constexpr int kProfLaunches = 64;
struct Prof {
  bool on;
  long long t0, acc[3];
  __device__ __forceinline__ explicit Prof(const ChainP &p) : on(p.prof != nullptr), t0(0), acc{0, 0, 0} {}
  __device__ __forceinline__ long long now() const { return on ? clock64() : 0; }
  __device__ __forceinline__ void start() { t0 = now(); }
  __device__ __forceinline__ void stop(int k) { if (on) acc[k] += clock64() - t0; }
  __device__ __forceinline__ void lap(const ChainP &p, int slot) {        // time since start(), straight to the sums
    if (on) atomicAdd(p.prof + slot, static_cast<unsigned long long>(clock64() - t0));
  }
  __device__ __forceinline__ void flush(const ChainP &p, int slot0, int n, long long total) {
    if (!on) return;
    for (int k = 0; k < n; ++k) atomicAdd(p.prof + slot0 + k, static_cast<unsigned long long>(acc[k]));
    atomicAdd(p.prof + slot0 + n, static_cast<unsigned long long>(total));
  }
};

// First column of tile `tile`.  With a column window the launch covers columns [win_off, win_off + win_cols) of every
// cloud only (tiles never straddle a window): tile t is the t-th tile of that sub-range, at its real column.
template <int NT>
__device__ __forceinline__ long long tile_col0(const ChainP &p, int tile) {
  const long long v = static_cast<long long>(tile) * NT;
  if (p.win_cols == 0) return v;
  const long long cloud = v / p.win_cols;
  return cloud * p.cols_per_cloud + p.win_off + (v - cloud * p.win_cols);
}

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
// try_wait parks the thread in hardware until the phase completes or the suspend-time hint expires; with the default
// (short) hint the producer, the MMA thread and 256 epilogue threads re-issue it continuously and took ~24 % of all
// issue slots of the chain kernels (ncu source counters, round 1).
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity, uint32_t hint) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity), "r"(hint) : "memory");
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
// The MMA warp runs its loops warp-uniformly (all 32 lanes take the same path, so the compiler keeps descriptors and
// addresses in uniform registers) and only the tcgen05 instructions themselves are predicated on the elected lane:
// with the loops inside `if (lane == 0)` ptxas wrapped every UTCHMMA in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop and
// a dependent descriptor chain, ~200 cycles per MMA issued against 32-64 cycles of tensor time (phase profile, r2).
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n" : "=r"(pred));
  return pred;
}
__device__ __forceinline__ void umma_commit(uint64_t *bar, uint32_t leader) {
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      "setp.ne.b32 q, %1, 0;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
      "}\n" ::"r"(smem_u32(bar)), "r"(leader) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc,
                                          uint32_t leader) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc), "r"(leader) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_shared_b32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void st_shared_v2(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void st_shared_u16(uint32_t addr, uint16_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(v) : "memory");
}

// bf16x2 pack: low half = bf16(a), high half = bf16(b)  (one cvt.rn.bf16x2.f32)
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t *>(&h);
}
// (hi, lo) split of two floats: H = {bf16(a), bf16(b)}, L = {bf16(a - hi_a), bf16(b - hi_b)}
__device__ __forceinline__ void split2(float a, float b, uint32_t &H, uint32_t &L) {
  H = pack_bf16(a, b);
  const float ha = __uint_as_float(H << 16), hb = __uint_as_float(H & 0xFFFF0000u);
  L = pack_bf16(a - ha, b - hb);
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

// Activation tile of one K-atom (64 channels): a "hi" and a "lo" part, each [NT rows x 128 B]
// with the 16-byte chunks XOR-swizzled by (row & 7).  Byte offset of channel e (0..63) of `row`
// inside a part, and of part (atom, lo?) inside the buffer:
__device__ __forceinline__ uint32_t part_off(int row, int e) {
  return static_cast<uint32_t>((row >> 3) * 1024 + (row & 7) * 128 + ((((e >> 3) ^ (row & 7)) << 4) | ((e & 7) << 1)));
}
template <int NT>
__device__ __forceinline__ uint32_t part_base(int atom, int lo) { return static_cast<uint32_t>((2 * atom + lo) * (NT * 128)); }

template <int NT>
__device__ __forceinline__ void store_scalar(uint32_t buf, int row, int ch, float v) {
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
  const uint32_t o = part_base<NT>(ch >> 6, 0) + part_off(row, ch & 63);
  st_shared_u16(buf + o, *reinterpret_cast<const uint16_t *>(&h));
  st_shared_u16(buf + o + NT * 128, *reinterpret_cast<const uint16_t *>(&l));
}

template <int NT>
__device__ __forceinline__ void store_quad(uint32_t buf, int row, int c4, float4 v) {
  uint32_t H01, L01, H23, L23;
  split2(v.x, v.y, H01, L01);
  split2(v.z, v.w, H23, L23);
  const uint32_t o = part_base<NT>(c4 >> 4, 0) + part_off(row, (c4 & 15) * 4);
  st_shared_v2(buf + o, H01, H23);
  st_shared_v2(buf + o + NT * 128, L01, L23);
}

// ---- input tile builder (256 worker threads) --------------------------------------------------
// Phase A: one thread per row resolves the row's source indices (one dependent-load chain for
// the whole tile) and writes the 3 recentred coordinates / the zero padding.  Phase B: one warp
// per row copies (or interpolates) the feature rows, 8 rows in flight.
template <int NT, int NW>
__device__ void load_tile(const ChainP &p, uint32_t buf, long long col0, int w8, int lane, int *s_arow,
                          int *s_brow, float *s_w, int bar_id) {
  const int t = w8 * 32 + lane;
  const int cin = p.l0_w != nullptr ? p.l0_cout
                                     : p.a_ch + (p.in_mode == CPFN_MLP_IN_GROUP ? (p.xyz_w != nullptr ? 0 : 3)
                                                                                : (p.in_mode == CPFN_MLP_IN_INTERP ? p.b_ch : 0));
  const int k16 = p.L[0].ksteps * 16;
  if (t < NT) {
    const long long col = col0 + t;
    const bool valid = col < p.cols;
    const long long c = valid ? col : 0;
    const long long cloud = c / p.cols_per_cloud;
    int arow = static_cast<int>(c);
    int pad_from = cin;
    if (p.in_mode == CPFN_MLP_IN_GROUP && p.xyz_w != nullptr) {
      arow = static_cast<int>(cloud * p.a_rows + __ldg(p.idx + c));       // positions enter in layer 0's epilogue
    } else if (p.in_mode == CPFN_MLP_IN_GROUP) {
      arow = static_cast<int>(cloud * p.a_rows + __ldg(p.idx + c));
      // recentred position (pointset_abstraction.py:62-63): xyz[idx] - centre
      const float *a = p.xyz + static_cast<long long>(arow) * 3;
      const float *ce = p.centers + (c / p.group_k) * 3;
      float d3[3];
#pragma unroll
      for (int q = 0; q < 3; ++q) d3[q] = valid ? __fsub_rn(__ldg(a + q), __ldg(ce + q)) : 0.f;
      if (p.l0_w != nullptr && NW == 4 && p.l0_cout == 64) {
        // 64-channel first layer, four loader warps: the rows only leave their recentred positions here; the layer
        // itself runs below with the weights in registers (lane = 4 channels)
        reinterpret_cast<float4 *>(s_brow)[t] = make_float4(d3[0], d3[1], d3[2], 0.f);
      } else if (p.l0_w != nullptr) {
        // first shared-MLP layer (3 input channels) in fp32 on the CUDA cores: relu(W0 d + b0); a K = 3
        // contraction would waste 13/16 of a tensor-core pass and a whole MMA / epilogue round trip
        for (int c4 = 0; c4 < (p.l0_cout >> 2); ++c4) {
          // 4 channels = 12 consecutive weights (three 16-byte loads, identical for the whole warp) + 4 biases
          const float4 wa = __ldg(reinterpret_cast<const float4 *>(p.l0_w) + c4 * 3);
          const float4 wb = __ldg(reinterpret_cast<const float4 *>(p.l0_w) + c4 * 3 + 1);
          const float4 wc = __ldg(reinterpret_cast<const float4 *>(p.l0_w) + c4 * 3 + 2);
          const float4 bb = __ldg(reinterpret_cast<const float4 *>(p.l0_b) + c4);
          float o[4];
          o[0] = fmaxf(fmaf(wa.z, d3[2], fmaf(wa.y, d3[1], fmaf(wa.x, d3[0], bb.x))), 0.f);
          o[1] = fmaxf(fmaf(wb.y, d3[2], fmaf(wb.x, d3[1], fmaf(wa.w, d3[0], bb.y))), 0.f);
          o[2] = fmaxf(fmaf(wc.x, d3[2], fmaf(wb.w, d3[1], fmaf(wb.z, d3[0], bb.z))), 0.f);
          o[3] = fmaxf(fmaf(wc.w, d3[2], fmaf(wc.z, d3[1], fmaf(wc.y, d3[0], bb.w))), 0.f);
          store_quad<NT>(buf, t, c4, valid ? make_float4(o[0], o[1], o[2], o[3]) : make_float4(0.f, 0.f, 0.f, 0.f));
        }
      } else {
#pragma unroll
        for (int q = 0; q < 3; ++q) store_scalar<NT>(buf, t, p.a_ch + q, d3[q]);
      }
    } else if (p.in_mode == CPFN_MLP_IN_INTERP) {
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        s_brow[t * 4 + q] = valid ? static_cast<int>(cloud * p.b_rows + __ldg(p.idx + c * 3 + q)) : 0;
        s_w[t * 4 + q] = valid ? __ldg(p.nn_w + c * 3 + q) : 0.f;
      }
    }
    s_arow[t] = valid ? arow : -1;
    for (int ch = pad_from; ch < k16; ++ch) store_scalar<NT>(buf, t, ch, 0.f);   // K padding must be finite zeros
  }
  asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(NW * 32) : "memory");
  if (p.in_mode == CPFN_MLP_IN_GROUP && p.l0_w != nullptr && NW == 4 && p.l0_cout == 64) {
    // relu(W0 d + b0) for the tile, lane = 4 output channels whose 12 weights + 4 biases stay in registers (the per-row
    // version re-loaded all 256 of them for every row: 64 uniform loads per row).  A warp covers rows
    // (lane >> 4) + 2 * warp + 8 * it: (row & 7) is fixed per thread, so the swizzled store offset is a constant plus
    // 1024 bytes per step and the 16 steps unroll into immediate offsets.  Same arithmetic, bit for bit, as the loop above.
    const int q = lane & 15, r7 = (lane >> 4) + 2 * w8;
    const float4 wa = __ldg(reinterpret_cast<const float4 *>(p.l0_w) + q * 3);
    const float4 wb = __ldg(reinterpret_cast<const float4 *>(p.l0_w) + q * 3 + 1);
    const float4 wc = __ldg(reinterpret_cast<const float4 *>(p.l0_w) + q * 3 + 2);
    const float4 bb = __ldg(reinterpret_cast<const float4 *>(p.l0_b) + q);
    const uint32_t o_hi = buf + static_cast<uint32_t>(r7 * 128) + static_cast<uint32_t>((((q >> 1) ^ r7) << 4) | ((q & 1) << 3));
    const float4 *sd = reinterpret_cast<const float4 *>(s_brow) + r7;
#pragma unroll
    for (int it = 0; it < NT / 8; ++it) {
      const float4 d = sd[8 * it];
      const float o0 = fmaxf(fmaf(wa.z, d.z, fmaf(wa.y, d.y, fmaf(wa.x, d.x, bb.x))), 0.f);
      const float o1 = fmaxf(fmaf(wb.y, d.z, fmaf(wb.x, d.y, fmaf(wa.w, d.x, bb.y))), 0.f);
      const float o2 = fmaxf(fmaf(wc.x, d.z, fmaf(wb.w, d.y, fmaf(wb.z, d.x, bb.z))), 0.f);
      const float o3 = fmaxf(fmaf(wc.w, d.z, fmaf(wc.z, d.y, fmaf(wc.y, d.x, bb.w))), 0.f);
      uint32_t H01, L01, H23, L23;
      split2(o0, o1, H01, L01);
      split2(o2, o3, H23, L23);
      st_shared_v2(o_hi + it * 1024, H01, H23);
      st_shared_v2(o_hi + it * 1024 + NT * 128, L01, L23);
    }
    return;
  }
  constexpr int RB = 8;
  const int nA4 = p.a_ch >> 2;
  if (nA4 > 0) {
    for (int r0 = w8; r0 < NT; r0 += NW * RB) {
      int arow[RB];
#pragma unroll
      for (int u = 0; u < RB; ++u) arow[u] = (r0 + NW * u < NT) ? s_arow[r0 + NW * u] : -1;
      for (int c4 = lane; c4 < nA4; c4 += 32) {
        float4 v[RB];
#pragma unroll
        for (int u = 0; u < RB; ++u)
          v[u] = arow[u] >= 0 ? __ldg(reinterpret_cast<const float4 *>(p.a_src + static_cast<long long>(arow[u]) * p.a_ch) + c4)
                              : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < RB; ++u)
          if (r0 + NW * u < NT) store_quad<NT>(buf, r0 + NW * u, c4, v[u]);
      }
    }
  }
  if (p.in_mode == CPFN_MLP_IN_INTERP && NT == 128 && NW == 8 && p.a_ch == 0 && p.b_ch == 128) {
    // Interpolation-only input of 128 channels (the FP3 + heads chain): lane = 4 channels of rows w8, w8 + 8, ... --
    // all of them have (row & 7) == w8, so the swizzled store offset is one constant plus 1024 bytes per row step, and
    // the per-row source rows / weights come as two 16-byte shared-memory records.  Rows past the end of the data
    // read row 0 with zero weights (finite values; a garbage row only affects its own, never stored, output row).
    constexpr int RI = 4;
    const float4 *src = reinterpret_cast<const float4 *>(p.b_src) + lane;
    float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.in_bias != nullptr) bb = __ldg(reinterpret_cast<const float4 *>(p.in_bias) + lane);
    const bool act_in = p.in_bias != nullptr;
    const uint32_t o_hi = buf + part_base<NT>(lane >> 4, 0) + static_cast<uint32_t>(w8 * 128) +
                          static_cast<uint32_t>(((((lane & 15) >> 1) ^ w8) << 4) | ((lane & 1) << 3));
#pragma unroll 1
    for (int j0 = 0; j0 < NT / 8; j0 += RI) {
      float4 f[RI][3], w[RI];
#pragma unroll
      for (int u = 0; u < RI; ++u) {
        const int r = w8 + 8 * (j0 + u);
        const int4 rows = *reinterpret_cast<const int4 *>(s_brow + r * 4);
        w[u] = *reinterpret_cast<const float4 *>(s_w + r * 4);
        f[u][0] = __ldg(src + static_cast<size_t>(static_cast<unsigned>(rows.x)) * 32);
        f[u][1] = __ldg(src + static_cast<size_t>(static_cast<unsigned>(rows.y)) * 32);
        f[u][2] = __ldg(src + static_cast<size_t>(static_cast<unsigned>(rows.z)) * 32);
      }
#pragma unroll
      for (int u = 0; u < RI; ++u) {
        float4 o;
        o.x = __fmaf_rn(f[u][2].x, w[u].z, __fmaf_rn(f[u][0].x, w[u].x, __fmul_rn(f[u][1].x, w[u].y)));
        o.y = __fmaf_rn(f[u][2].y, w[u].z, __fmaf_rn(f[u][0].y, w[u].x, __fmul_rn(f[u][1].y, w[u].y)));
        o.z = __fmaf_rn(f[u][2].z, w[u].z, __fmaf_rn(f[u][0].z, w[u].x, __fmul_rn(f[u][1].z, w[u].y)));
        o.w = __fmaf_rn(f[u][2].w, w[u].z, __fmaf_rn(f[u][0].w, w[u].x, __fmul_rn(f[u][1].w, w[u].y)));
        if (act_in) {
          o.x = fmaxf(o.x + bb.x, 0.f); o.y = fmaxf(o.y + bb.y, 0.f);
          o.z = fmaxf(o.z + bb.z, 0.f); o.w = fmaxf(o.w + bb.w, 0.f);
        }
        uint32_t H01, L01, H23, L23;
        split2(o.x, o.y, H01, L01);
        split2(o.z, o.w, H23, L23);
        const uint32_t a = o_hi + static_cast<uint32_t>(j0 + u) * 1024;
        st_shared_v2(a, H01, H23);
        st_shared_v2(a + NT * 128, L01, L23);
      }
    }
  } else if (p.in_mode == CPFN_MLP_IN_INTERP) {
    // three_weighted_sum (interpolate_gpu.cu:98-99 as compiled): fma(p3,w3, fma(p1,w1, p2*w2))
    constexpr int RI = 4;
    const int nB4 = p.b_ch >> 2, a4 = p.a_ch >> 2;
    for (int r0 = w8; r0 < NT; r0 += NW * RI) {
      for (int c4 = lane; c4 < nB4; c4 += 32) {
        float4 f[RI][3];
        float w[RI][3];
        bool ok[RI];
#pragma unroll
        for (int u = 0; u < RI; ++u) {
          const int r = r0 + NW * u;
          ok[u] = r < NT && s_arow[r < NT ? r : 0] >= 0;
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            w[u][q] = ok[u] ? s_w[r * 4 + q] : 0.f;
            f[u][q] = ok[u] ? __ldg(reinterpret_cast<const float4 *>(p.b_src + static_cast<long long>(s_brow[r * 4 + q]) * p.b_ch) + c4)
                            : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
#pragma unroll
        for (int u = 0; u < RI; ++u) {
          const int r = r0 + NW * u;
          if (r >= NT) continue;
          float4 o;
          o.x = __fmaf_rn(f[u][2].x, w[u][2], __fmaf_rn(f[u][0].x, w[u][0], __fmul_rn(f[u][1].x, w[u][1])));
          o.y = __fmaf_rn(f[u][2].y, w[u][2], __fmaf_rn(f[u][0].y, w[u][0], __fmul_rn(f[u][1].y, w[u][1])));
          o.z = __fmaf_rn(f[u][2].z, w[u][2], __fmaf_rn(f[u][0].z, w[u][0], __fmul_rn(f[u][1].z, w[u][1])));
          o.w = __fmaf_rn(f[u][2].w, w[u][2], __fmaf_rn(f[u][0].w, w[u][0], __fmul_rn(f[u][1].w, w[u][1])));
          if (p.in_bias != nullptr) {      // the layer in front of the interpolation: its bias and ReLU
            const float4 bb = __ldg(reinterpret_cast<const float4 *>(p.in_bias) + c4);
            o.x = ok[u] ? fmaxf(o.x + bb.x, 0.f) : 0.f; o.y = ok[u] ? fmaxf(o.y + bb.y, 0.f) : 0.f;
            o.z = ok[u] ? fmaxf(o.z + bb.z, 0.f) : 0.f; o.w = ok[u] ? fmaxf(o.w + bb.w, 0.f) : 0.f;
          }
          store_quad<NT>(buf, r, a4 + c4, o);
        }
      }
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
  int *s_arow = reinterpret_cast<int *>(bars + 32);          // [128]
  int *s_brow = s_arow + 128;                                // [128][4] (16-byte records: three source rows)
  float *s_w = reinterpret_cast<float *>(s_brow + 512);      // [128][4] (three weights)
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
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // warp-uniform for the compiler
  const int wave_max = p.tmem_cols / NT;

  if (warp == 0) {
    // ===== weight producer: the packed blob is consumed strictly in order, once per tile =====
    if (lane == 0) {
      Prof pf(p);
      const long long pf_begin = pf.now();
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const int nblk = p.split_cout ? p.L[0].cin_atoms * 2 : p.total_blocks;
        const size_t blk0 = p.split_cout ? static_cast<size_t>(blockIdx.y) * nblk : 0;
        for (int blk = 0; blk < nblk; ++blk) {
          pf.start();
          mbar_wait(empty + stage, phase ^ 1, p.wait_hint);
          pf.stop(0);
          mbar_arrive_expect_tx(full + stage, kStageBytes);
          bulk_g2s(ring + stage * kStageBytes, p.weights + (blk0 + blk) * kStageBytes, kStageBytes,
                   full + stage);
          if (++stage == p.nstage) { stage = 0; phase ^= 1; }
        }
      }
      pf.flush(p, 3, 1, pf.now() - pf_begin);
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread): per K-atom  D += Whi*Ahi + Whi*Alo  then  D += Wlo*Ahi =====
    {
      const uint32_t leader = elect_one();
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(NT >> 3) << 17) |
                                 (static_cast<uint32_t>(128 >> 4) << 24);
      Prof pf(p);
      const long long pf_begin = pf.now();
      int stage = 0;
      uint32_t phase = 0, act_phase = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        for (int l = 0; l < p.n_layers; ++l) {
          const LayerP &L = p.L[l];
          const uint32_t in_buf = smem_u32(act0);   // in-place: a layer's output overwrites its (fully consumed) input
          const int n_chunks = p.split_cout ? 1 : L.cout_chunks;
          for (int m0 = 0; m0 < n_chunks; m0 += wave_max) {
            const int mc = min(wave_max, n_chunks - m0);
            pf.start();
            mbar_wait(act_ready, act_phase, p.wait_hint);
            pf.stop(0);
            act_phase ^= 1;
            tc_fence_after();
            for (int m = 0; m < mc; ++m) {
              const uint32_t d_tmem = tmem_base + m * NT;
              for (int j = 0; j < L.cin_atoms; ++j) {
                const int ks = min(4, L.ksteps - 4 * j);
                const uint32_t a_hi = in_buf + part_base<NT>(j, 0), a_lo = a_hi + NT * 128;
                pf.start();
                mbar_wait(full + stage, phase, p.wait_hint);            // W_hi block
                pf.stop(1);
                tc_fence_after();
                // descriptors of K step kk = those of step 0 + 2 (32 bytes >> 4) in the 14-bit start-address field
                const uint64_t d_ahi = make_desc(a_hi), d_alo = make_desc(a_lo);
                uint64_t d_w = make_desc(smem_u32(ring + stage * kStageBytes));
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                  if (kk < ks) {
                    umma_bf16(d_tmem, d_w + 2 * kk, d_ahi + 2 * kk, idesc, (j | kk) != 0 ? 1u : 0u, leader);
                    umma_bf16(d_tmem, d_w + 2 * kk, d_alo + 2 * kk, idesc, 1u, leader);
                  }
                umma_commit(empty + stage, leader);                // frees the stage when these MMAs retire
                if (++stage == p.nstage) { stage = 0; phase ^= 1; }
                pf.start();
                mbar_wait(full + stage, phase, p.wait_hint);            // W_lo block
                pf.stop(1);
                tc_fence_after();
                d_w = make_desc(smem_u32(ring + stage * kStageBytes));
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                  if (kk < ks) umma_bf16(d_tmem, d_w + 2 * kk, d_ahi + 2 * kk, idesc, 1u, leader);
                umma_commit(empty + stage, leader);
                if (++stage == p.nstage) { stage = 0; phase ^= 1; }
              }
            }
            umma_commit(acc_full, leader);                         // accumulators of this wave complete
          }
        }
      }
      if (leader) pf.flush(p, 0, 2, pf.now() - pf_begin);
    }
  } else {
    // ===== workers: build the input tile, then the epilogue of every layer =====
    const int w8 = warp - 2;                          // 0..7
    const int wq = warp & 3;                          // TMEM lane quarter this warp may access
    const int half = (w8 >> 2);                       // which half of the tile's columns
    const int par = lane & 1;
    const int row_in_chunk = wq * 32 + lane;
    const int e_pair = ((wq & 1) * 32 + lane) & ~1;   // channel (within the 64-wide atom) of the pair's even lane
    uint32_t swz[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) swz[k] = part_off(2 * k + par, e_pair);
    const uint32_t sel_send = par ? 0x5410u : 0x7632u, sel_keep = par ? 0x7632u : 0x5410u;
    constexpr int HC = NT / 2;                        // columns per worker warp
    uint32_t acc_phase = 0;
    Prof pf(p);
    pf.on = pf.on && threadIdx.x == 64;
    const long long pf_begin = pf.now();
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      const long long col0 = tile_col0<NT>(p, tile);
      const long long cloud = col0 / p.cols_per_cloud;
      const long long n_in_cloud = col0 - cloud * p.cols_per_cloud;
      pf.start();
      load_tile<NT, 8>(p, smem_u32(act0), col0, w8, lane, s_arow, s_brow, s_w, 1);
      fence_proxy_async();
      mbar_arrive(act_ready);
      pf.stop(0);
      pf.acc[2] += 1;
      for (int l = 0; l < p.n_layers; ++l) {
        const LayerP &L = p.L[l];
        const bool last = (l == p.n_layers - 1);
        const uint32_t out_buf = smem_u32(act0);
        const int cout_pad = L.cout_chunks * 128;
        const bool slow = (L.mask_bits != nullptr) || (L.out_cm != nullptr);
        const int n_chunks = p.split_cout ? 1 : L.cout_chunks;
        const int chunk_base = p.split_cout ? static_cast<int>(blockIdx.y) : 0;
        for (int m0 = 0; m0 < n_chunks; m0 += wave_max) {
          const int mc = min(wave_max, n_chunks - m0);
          pf.start();
          mbar_wait(acc_full, acc_phase, p.wait_hint);
          pf.stop(1);
          pf.start();
          acc_phase ^= 1;
          tc_fence_after();
          for (int m = 0; m < mc; ++m) {
            const int ch = (chunk_base + m0 + m) * 128 + row_in_chunk;
            const bool ch_real = ch < L.cout;
            const float bias = __ldg(L.bias + (L.bias_per_cloud ? cloud * cout_pad : 0) + ch);
            const bool to_smem = !last && (ch & ~1) < L.next_k16;
            const uint32_t o_hi = out_buf + part_base<NT>(ch >> 6, 0);
            const size_t cm_base = (static_cast<size_t>(cloud) * L.cout + ch) * p.cols_per_cloud + n_in_cloud;
            // Max-pool of the last layer: bias and ReLU are monotone and per channel, so the raw accumulators are pooled and
            // relu(max + bias) is formed once per group (identical bits, a quarter of the instructions per element).
            const bool pool_last = last && p.out_mode == CPFN_MLP_OUT_POOL;
            const bool tile_full = col0 + NT <= p.cols;
            float pool = -INFINITY;
#pragma unroll 1
            for (int cb = 0; cb < HC; cb += 16) {
              const int c = half * HC + cb;            // first column of this batch inside the tile
              uint32_t r[16];
              tmem_ld16(tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + m * NT + c, r);
              if (pool_last) {
                if (tile_full) {
#pragma unroll
                  for (int i = 0; i < 16; ++i) pool = fmaxf(pool, __uint_as_float(r[i]));
                } else {
#pragma unroll
                  for (int i = 0; i < 16; ++i) pool = fmaxf(pool, (col0 + c + i < p.cols) ? __uint_as_float(r[i]) : -INFINITY);
                }
                if (p.pool_g <= HC && ((c + 16) % p.pool_g) == 0) {
                  if (ch_real && col0 + c + 16 - p.pool_g < p.cols) p.out[((col0 + c) / p.pool_g) * p.ldo + ch] = fmaxf(pool + bias, 0.f);
                  pool = -INFINITY;
                }
                continue;
              }
              float v[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                v[i] = __uint_as_float(r[i]) + bias;
                if (L.relu) v[i] = fmaxf(v[i], 0.f);
              }
              if (slow && ch_real) {                   // dropout mask / channel-major copy (fc1 layer)
#pragma unroll
                for (int i4 = 0; i4 < 4; ++i4) {
                  if (col0 + c + i4 * 4 < p.cols) {    // tiles are cloud-aligned here: whole quads are valid
                    if (L.mask_bits != nullptr) {
                      const uint32_t *mb = L.mask_bits + (col0 + c + i4 * 4) * L.mask_words + (ch >> 5);
#pragma unroll
                      for (int q = 0; q < 4; ++q)
                        v[i4 * 4 + q] = ((__ldg(mb + q * L.mask_words) >> (ch & 31)) & 1u) ? v[i4 * 4 + q] * L.mask_scale : 0.f;
                    }
                    if (L.out_cm != nullptr)
                      reinterpret_cast<float4 *>(L.out_cm + cm_base + c)[i4] =
                          make_float4(v[i4 * 4], v[i4 * 4 + 1], v[i4 * 4 + 2], v[i4 * 4 + 3]);
                  }
                }
              }
              if (!last) {
                // next layer's operand: (hi, lo) bf16 split, channel pairs packed into 32-bit words
                const uint32_t rowgrp = o_hi + static_cast<uint32_t>(c >> 3) * 1024;
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                  uint32_t H, Lw;
                  split2(v[i], v[i + 1], H, Lw);
                  const uint32_t send = __byte_perm(H, Lw, sel_send);
                  const uint32_t keep = __byte_perm(H, Lw, sel_keep);
                  const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, 1);
                  const uint32_t a = par ? recv : keep, b = par ? keep : recv;
                  if (to_smem) {
                    const uint32_t addr = rowgrp + swz[(i >> 1) & 3] + (i >> 3) * 1024;
                    st_shared_b32(addr, __byte_perm(a, b, 0x5410));
                    st_shared_b32(addr + NT * 128, __byte_perm(a, b, 0x7632));
                  }
                }
              } else if (p.out_mode == CPFN_MLP_OUT_ROWS) {
                if (ch_real) {
#pragma unroll
                  for (int i = 0; i < 16; ++i)
                    if (col0 + c + i < p.cols) p.out[(col0 + c + i) * p.ldo + ch] = v[i];
                }
              }
            }
            if (pool_last && p.pool_g > HC && ch_real && col0 + half * HC < p.cols)
              atomicMax(reinterpret_cast<int *>(p.out + ((col0 + half * HC) / p.pool_g) * p.ldo + ch),
                        __float_as_int(fmaxf(pool + bias, 0.f)));
          }
          tc_fence_before();
          const bool final_wave = last && (m0 + wave_max >= n_chunks);
          if (!final_wave) {
            fence_proxy_async();
            mbar_arrive(act_ready);
          }
          pf.lap(p, 10 + min(l, 5));
        }
      }
    }
    if (pf.on) {
      const long long tiles = pf.acc[2];
      pf.flush(p, 5, 2, pf.now() - pf_begin);
      atomicAdd(p.prof + 8, 1ull);
      atomicAdd(p.prof + 9, static_cast<unsigned long long>(tiles));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

// =================================================================================================
// Points-as-M variant (tiles of 128 columns): D[point, channel] = A[point, k] * W[channel, k]^T.
// The activation tile is the A operand (M = 128 rows = TMEM lanes), a weight block is the B operand
// (N = up to 128 output channels = TMEM columns).  An epilogue thread owns one POINT and reads 16
// consecutive channels per tcgen05.ld, so the (hi, lo) split packs 8 channels into one 16-byte
// shared-memory store with no cross-lane traffic (about half the instructions per element of the
// channels-as-M kernel above, no padding of 64-wide layers to 128), the channel-major copy and the
// dropout mask are coalesced across the warp.  A pooled LAST layer swaps the operand roles (see epi_pool_cols);
// ragged tiles fall back to one REDUX per channel + atomicMax.
// =================================================================================================
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// ---- specialised epilogue loops of the points-as-M kernels (thread = point) ---------------------------------
// The generic per-chunk body below serves every combination of (ReLU, dropout bits, channel-major copy, rows / pooled /
// next-layer output); each extra uniform branch in it costs every layer (the 128 -> 128 layers of the heads chain went
// from 12.7 to 17 kilo-cycles per CTA as cases were added), so the combinations the network actually runs get their
// own straight-line loops.  `taddr` = TMEM address of the thread's lane quarter at column 0 of the chunk's accumulator.

// bias + (ReLU) + (hi, lo) split -> the next layer's operand tile, channels [chunk0 + cb0, chunk0 + cb1)
template <bool RELU>
__device__ __forceinline__ void epi_next(uint32_t taddr, const float *__restrict__ bias, int chunk0, int cb0, int cb1,
                                         int next_k16, uint32_t out_buf, uint32_t row_base, int r7) {
#pragma unroll 1
  for (int cb = cb0; cb < cb1; cb += 16) {
    const int ch0 = chunk0 + cb;
    uint32_t r[16];
    float4 b4[4];
#pragma unroll
    for (int i4 = 0; i4 < 4; ++i4) b4[i4] = __ldg(reinterpret_cast<const float4 *>(bias + ch0) + i4);
    tmem_ld16(taddr + cb, r);
    if (ch0 >= next_k16) continue;
    float v[16];
#pragma unroll
    for (int i4 = 0; i4 < 4; ++i4) {
      v[i4 * 4] = __uint_as_float(r[i4 * 4]) + b4[i4].x;
      v[i4 * 4 + 1] = __uint_as_float(r[i4 * 4 + 1]) + b4[i4].y;
      v[i4 * 4 + 2] = __uint_as_float(r[i4 * 4 + 2]) + b4[i4].z;
      v[i4 * 4 + 3] = __uint_as_float(r[i4 * 4 + 3]) + b4[i4].w;
    }
    if (RELU) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
    }
    uint32_t H[8], Lo[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) split2(v[2 * i], v[2 * i + 1], H[i], Lo[i]);
    const uint32_t base = out_buf + part_base<128>(ch0 >> 6, 0) + row_base;
    const int c8 = (ch0 & 63) >> 3;
    const uint32_t a0 = base + (((c8) ^ r7) << 4), a1 = base + (((c8 + 1) ^ r7) << 4);
    st_shared_v4(a0, H[0], H[1], H[2], H[3]);
    st_shared_v4(a1, H[4], H[5], H[6], H[7]);
    st_shared_v4(a0 + 128 * 128, Lo[0], Lo[1], Lo[2], Lo[3]);
    st_shared_v4(a1 + 128 * 128, Lo[4], Lo[5], Lo[6], Lo[7]);
  }
}

// the same with layer 0's position term instead of a plain bias: rows (wx, wy, wz, bias) per channel
template <bool RELU>
__device__ __forceinline__ void epi_next_xyz(uint32_t taddr, const float4 *__restrict__ wb, float dx, float dy, float dz,
                                             int chunk0, int cb0, int cb1, int next_k16, uint32_t out_buf,
                                             uint32_t row_base, int r7) {
#pragma unroll 1
  for (int cb = cb0; cb < cb1; cb += 16) {
    const int ch0 = chunk0 + cb;
    uint32_t r[16];
    float4 w4[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) w4[i] = __ldg(wb + ch0 + i);
    tmem_ld16(taddr + cb, r);
    if (ch0 >= next_k16) continue;
    float v[16];
#pragma unroll
    for (int i4 = 0; i4 < 4; ++i4) {          // four rows at a time: 16 registers of weights in flight, not 64
      float4 nx[4];
      if (i4 < 3) {
#pragma unroll
        for (int i = 0; i < 4; ++i) nx[i] = __ldg(wb + ch0 + (i4 + 1) * 4 + i);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float t = fmaf(w4[i].z, dz, fmaf(w4[i].y, dy, fmaf(w4[i].x, dx, __uint_as_float(r[i4 * 4 + i]) + w4[i].w)));
        v[i4 * 4 + i] = RELU ? fmaxf(t, 0.f) : t;
      }
      if (i4 < 3) {
#pragma unroll
        for (int i = 0; i < 4; ++i) w4[i] = nx[i];
      }
    }
    uint32_t H[8], Lo[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) split2(v[2 * i], v[2 * i + 1], H[i], Lo[i]);
    const uint32_t base = out_buf + part_base<128>(ch0 >> 6, 0) + row_base;
    const int c8 = (ch0 & 63) >> 3;
    const uint32_t a0 = base + (((c8) ^ r7) << 4), a1 = base + (((c8 + 1) ^ r7) << 4);
    st_shared_v4(a0, H[0], H[1], H[2], H[3]);
    st_shared_v4(a1, H[4], H[5], H[6], H[7]);
    st_shared_v4(a0 + 128 * 128, Lo[0], Lo[1], Lo[2], Lo[3]);
    st_shared_v4(a1 + 128 * 128, Lo[4], Lo[5], Lo[6], Lo[7]);
  }
}

// Max-pool of the LAST layer.  Its MMAs run with the operand roles swapped (weights = A, M = 128 channels; the tile's
// 128 points = B, N = 128 columns), so the accumulator has channels on the TMEM lanes and points along the columns: a
// thread pools its channel over each group of `pool_g` consecutive columns in registers -- no cross-lane reduction, no
// shared memory, no atomics -- and, bias and ReLU being monotone and per channel, applies them once per group to the
// pooled RAW accumulator (max_cols relu(acc + b) == relu(max_cols(acc) + b) exactly).  Lanes = consecutive channels:
// the store is one coalesced line per warp.  taddr = the warp's lane quarter at column 0 of the chunk's accumulator.
__device__ __forceinline__ void epi_pool_cols(uint32_t taddr, int col_begin, int pool_g, float bias, bool relu,
                                              float *__restrict__ out /* group's row + channel */, bool ch_ok) {
  float mx = -INFINITY;
#pragma unroll 1
  for (int cb = 0; cb < pool_g; cb += 16) {
    uint32_t r[16];
    tmem_ld16(taddr + col_begin + cb, r);
    float m4[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      m4[i] = fmaxf(fmaxf(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1])),
                    fmaxf(__uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3])));
    mx = fmaxf(mx, fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])));
  }
  float v = mx + bias;
  if (relu) v = fmaxf(v, 0.f);
  if (ch_ok) *out = v;
}

// bias + (ReLU) -> this row of the staged output tile (row stride row_ls floats, see the copy-out in the kernel)
template <bool RELU>
__device__ __forceinline__ void epi_rows_staged(uint32_t taddr, const float *__restrict__ bias, int chunk0, int cb0, int cb1,
                                                float *s_row) {
#pragma unroll 1
  for (int cb = cb0; cb < cb1; cb += 16) {
    const int ch0 = chunk0 + cb;
    uint32_t r[16];
    float4 b4[4];
#pragma unroll
    for (int i4 = 0; i4 < 4; ++i4) b4[i4] = __ldg(reinterpret_cast<const float4 *>(bias + ch0) + i4);
    tmem_ld16(taddr + cb, r);
    float4 *dst = reinterpret_cast<float4 *>(s_row + ch0);
#pragma unroll
    for (int i4 = 0; i4 < 4; ++i4) {
      float4 o = make_float4(__uint_as_float(r[i4 * 4]) + b4[i4].x, __uint_as_float(r[i4 * 4 + 1]) + b4[i4].y,
                             __uint_as_float(r[i4 * 4 + 2]) + b4[i4].z, __uint_as_float(r[i4 * 4 + 3]) + b4[i4].w);
      if (RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
      dst[i4] = o;
    }
  }
}

// Single-tile form of the points-as-M kernel (8 epilogue warps, each lane quarter's two warps share a
// chunk's channels): used when one 128-column tile per CTA lets two CTAs share an SM but two sub-tiles
// would not (the FP3 + fc1 + heads chain: 64 KB of activations per tile).
__global__ void __launch_bounds__(kChainThreads, 2)
mlp_chain_pm1_kernel(const __grid_constant__ ChainP p) {
  constexpr int NT = 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t *ring = smem;
  uint8_t *act0 = ring + p.nstage * kStageBytes;
  uint8_t *act1 = act0 + p.act_bytes0;
  uint64_t *bars = reinterpret_cast<uint64_t *>(act1 + p.act_bytes1);
  uint64_t *full = bars, *empty = bars + kMaxStages, *act_ready = bars + 2 * kMaxStages,
           *acc_full = bars + 2 * kMaxStages + 1;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * kMaxStages + 2);
  int *s_arow = reinterpret_cast<int *>(bars + 32);
  int *s_brow = s_arow + 128;
  float *s_w = reinterpret_cast<float *>(s_brow + 512);
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
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // warp-uniform for the compiler
  const int wave_max = p.tmem_cols / 128;

  if (warp == 0) {
    if (lane == 0) {
      Prof pf(p);
      const long long pf_begin = pf.now();
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        for (int blk = 0; blk < p.total_blocks; ++blk) {
          pf.start();
          mbar_wait(empty + stage, phase ^ 1, p.wait_hint);
          pf.stop(0);
          mbar_arrive_expect_tx(full + stage, kStageBytes);
          bulk_g2s(ring + stage * kStageBytes, p.weights + static_cast<size_t>(blk) * kStageBytes, kStageBytes,
                   full + stage);
          if (++stage == p.nstage) { stage = 0; phase ^= 1; }
        }
      }
      pf.flush(p, 3, 1, pf.now() - pf_begin);
    }
  } else if (warp == 1) {
    {
      const uint32_t leader = elect_one();
      constexpr uint32_t idesc0 = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(128 >> 4) << 24);
      Prof pf(p);
      const long long pf_begin = pf.now();
      int stage = 0;
      uint32_t phase = 0, act_phase = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        for (int l = 0; l < p.n_layers; ++l) {
          const LayerP &L = p.L[l];
          const uint32_t in_buf = smem_u32(act0);   // in-place: a layer's output overwrites its (fully consumed) input
          const int cout16 = (L.cout + 15) & ~15;
          for (int m0 = 0; m0 < L.cout_chunks; m0 += wave_max) {
            const int mc = min(wave_max, L.cout_chunks - m0);
            pf.start();
            mbar_wait(act_ready, act_phase, p.wait_hint);
            pf.stop(0);
            act_phase ^= 1;
            tc_fence_after();
            for (int m = 0; m < mc; ++m) {
              const uint32_t d_tmem = tmem_base + m * 128;
              const int nch = min(128, cout16 - (m0 + m) * 128);
              const bool swap = p.pool_fast && l == p.n_layers - 1;        // pooled last layer: channels on the TMEM lanes
              const uint32_t idesc = idesc0 | (static_cast<uint32_t>((swap ? 128 : nch) >> 3) << 17);
              for (int j = 0; j < L.cin_atoms; ++j) {
                const int ks = min(4, L.ksteps - 4 * j);
                const uint32_t a_hi = in_buf + part_base<NT>(j, 0), a_lo = a_hi + NT * 128;
                pf.start();
                mbar_wait(full + stage, phase, p.wait_hint);            // W_hi block
                pf.stop(1);
                tc_fence_after();
                const uint64_t d_ahi = make_desc(a_hi), d_alo = make_desc(a_lo);
                uint64_t d_w = make_desc(smem_u32(ring + stage * kStageBytes));
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                  if (kk < ks) {
                    umma_bf16(d_tmem, swap ? d_w + 2 * kk : d_ahi + 2 * kk, swap ? d_ahi + 2 * kk : d_w + 2 * kk, idesc,
                              (j | kk) != 0 ? 1u : 0u, leader);
                    umma_bf16(d_tmem, swap ? d_w + 2 * kk : d_alo + 2 * kk, swap ? d_alo + 2 * kk : d_w + 2 * kk, idesc, 1u, leader);
                  }
                umma_commit(empty + stage, leader);
                if (++stage == p.nstage) { stage = 0; phase ^= 1; }
                pf.start();
                mbar_wait(full + stage, phase, p.wait_hint);            // W_lo block
                pf.stop(1);
                tc_fence_after();
                d_w = make_desc(smem_u32(ring + stage * kStageBytes));
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                  if (kk < ks)
                    umma_bf16(d_tmem, swap ? d_w + 2 * kk : d_ahi + 2 * kk, swap ? d_ahi + 2 * kk : d_w + 2 * kk, idesc, 1u, leader);
                umma_commit(empty + stage, leader);
                if (++stage == p.nstage) { stage = 0; phase ^= 1; }
              }
            }
            umma_commit(acc_full, leader);
          }
        }
      }
      if (leader) pf.flush(p, 0, 2, pf.now() - pf_begin);
    }
  } else {
    const int w8 = warp - 2;
    const int wq = warp & 3;                          // TMEM lane quarter = rows wq*32 .. wq*32+31 of the tile
    const int half = (w8 >> 2);                       // which 64-channel half of every 128-channel chunk
    const int row = wq * 32 + lane;
    const uint32_t row_base = static_cast<uint32_t>((row >> 3) * 1024 + (row & 7) * 128);
    const int r7 = row & 7;
    uint32_t acc_phase = 0;
    Prof pf(p);
    pf.on = pf.on && threadIdx.x == 64;
    const long long pf_begin = pf.now();
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      const long long col0 = tile_col0<NT>(p, tile);
      const long long col = col0 + row;
      const bool row_ok = col < p.cols;
      // thread = point: the cloud of THIS row (a tile may straddle two clouds when N is not a multiple of 128)
      const long long cloud = (row_ok ? col : col0) / p.cols_per_cloud;
      const long long n_in_cloud = (row_ok ? col : col0) - cloud * p.cols_per_cloud - row;
      pf.start();
      load_tile<NT, 8>(p, smem_u32(act0), col0, w8, lane, s_arow, s_brow, s_w, 1);
      fence_proxy_async();
      mbar_arrive(act_ready);
      pf.stop(0);
      pf.acc[2] += 1;
      // position columns of layer 0 as an epilogue term: this row's recentred position (pointset_abstraction.py:62-63)
      float d3x = 0.f, d3y = 0.f, d3z = 0.f;
      if (p.xyz_w != nullptr && row_ok) {
        const float *a = p.xyz + (cloud * p.a_rows + __ldg(p.idx + col)) * 3;
        const float *ce = p.centers + (col / p.group_k) * 3;
        d3x = __fsub_rn(__ldg(a), __ldg(ce)); d3y = __fsub_rn(__ldg(a + 1), __ldg(ce + 1)); d3z = __fsub_rn(__ldg(a + 2), __ldg(ce + 2));
      }
      for (int l = 0; l < p.n_layers; ++l) {
        const LayerP &L = p.L[l];
        const bool last = (l == p.n_layers - 1);
        const bool xyz_term = l == 0 && p.xyz_w != nullptr;
        const uint32_t out_buf = smem_u32(act0);
        const int cout_pad = L.cout_chunks * 128, cout16 = (L.cout + 15) & ~15;
        const bool slow = (L.mask_bits != nullptr) || (L.out_cm != nullptr);
        const int row_ls = cout16 + 4;                  // staged output row stride (floats): conflict-free float4 stores
        const bool stage_rows = last && p.out_mode == CPFN_MLP_OUT_ROWS && L.cout_chunks <= wave_max &&
                                128 * row_ls * 4 <= p.act_bytes0;
        const float *bias_base = L.bias + (L.bias_per_cloud ? cloud * cout_pad : 0);
        uint4 mbits = make_uint4(~0u, ~0u, ~0u, ~0u);     // this point's keep bits (thread = point)
        if (L.mask_bits != nullptr && row_ok) {
          const uint32_t *mb = L.mask_bits + col * L.mask_words;
          mbits.x = __ldg(mb);
          if (L.mask_words > 1) mbits.y = __ldg(mb + 1);
          if (L.mask_words > 2) mbits.z = __ldg(mb + 2);
          if (L.mask_words > 3) mbits.w = __ldg(mb + 3);
        }
        for (int m0 = 0; m0 < L.cout_chunks; m0 += wave_max) {
          const int mc = min(wave_max, L.cout_chunks - m0);
          pf.start();
          mbar_wait(acc_full, acc_phase, p.wait_hint);
          pf.stop(1);
          pf.start();
          acc_phase ^= 1;
          tc_fence_after();
          for (int m = 0; m < mc; ++m) {
            const int chunk0 = (m0 + m) * 128;
            const int nch = min(128, cout16 - chunk0);
            const int split = ((nch / 16 + 1) / 2) * 16;     // the quarter's two warps share the chunk's channels
            const int cb0 = half ? split : 0, cb1 = half ? nch : split;
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + m * 128;
            // the combinations the network runs: straight-line loops (see epi_next); everything else: the generic body
            if (last && p.pool_fast) {
              // (chunk, pooling group) items, shared between the quarter's two warps
              const int n_groups = 128 / p.pool_g;
              const int ch = chunk0 + wq * 32 + lane;
              const bool ch_ok = ch < L.cout;
              for (int gi = 0; gi < n_groups; ++gi) {
                if ((((m0 + m) * n_groups + gi) & 1) != half) continue;
                const long long gcol = col0 + static_cast<long long>(gi) * p.pool_g;
                const float bias = ch_ok ? __ldg(L.bias + (L.bias_per_cloud ? (gcol / p.cols_per_cloud) * cout_pad : 0) + ch) : 0.f;
                epi_pool_cols(taddr, gi * p.pool_g, p.pool_g, bias, L.relu != 0, p.out + (gcol / p.pool_g) * p.ldo + ch, ch_ok);
              }
              continue;
            }
            if (!last && !slow) {
              if (xyz_term) {
                if (L.relu) epi_next_xyz<true>(taddr, reinterpret_cast<const float4 *>(p.xyz_w), d3x, d3y, d3z, chunk0, cb0, cb1, L.next_k16, out_buf, row_base, r7);
                else epi_next_xyz<false>(taddr, reinterpret_cast<const float4 *>(p.xyz_w), d3x, d3y, d3z, chunk0, cb0, cb1, L.next_k16, out_buf, row_base, r7);
              } else {
                if (L.relu) epi_next<true>(taddr, bias_base, chunk0, cb0, cb1, L.next_k16, out_buf, row_base, r7);
                else epi_next<false>(taddr, bias_base, chunk0, cb0, cb1, L.next_k16, out_buf, row_base, r7);
              }
              continue;
            }
            if (stage_rows && !slow) {
              float *s_row = reinterpret_cast<float *>(act0) + row * row_ls;
              if (L.relu) epi_rows_staged<true>(taddr, bias_base, chunk0, cb0, cb1, s_row);
              else epi_rows_staged<false>(taddr, bias_base, chunk0, cb0, cb1, s_row);
              continue;
            }
#pragma unroll 1
            for (int cb = cb0; cb < cb1; cb += 16) {
              const int ch0 = chunk0 + cb;
              uint32_t r[16];
              float v[16];
              {
                float4 b4[4];                            // the bias loads are in flight while the accumulators arrive
#pragma unroll
                for (int i4 = 0; i4 < 4; ++i4) b4[i4] = __ldg(reinterpret_cast<const float4 *>(bias_base + ch0) + i4);
                tmem_ld16(taddr + cb, r);
#pragma unroll
                for (int i4 = 0; i4 < 4; ++i4) {
                  v[i4 * 4] = __uint_as_float(r[i4 * 4]) + b4[i4].x;
                  v[i4 * 4 + 1] = __uint_as_float(r[i4 * 4 + 1]) + b4[i4].y;
                  v[i4 * 4 + 2] = __uint_as_float(r[i4 * 4 + 2]) + b4[i4].z;
                  v[i4 * 4 + 3] = __uint_as_float(r[i4 * 4 + 3]) + b4[i4].w;
                }
              }
              if (L.relu) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
              }
              if (slow && row_ok) {                    // dropout mask / channel-major copy: coalesced over the warp
                const size_t o0 = (static_cast<size_t>(cloud) * L.cout + ch0) * p.cols_per_cloud + n_in_cloud + row;
                const size_t cs = static_cast<size_t>(p.cols_per_cloud);
                if (L.mask_bits != nullptr) {
                  const int mw = ch0 >> 5;
                  const uint32_t m16 = (mw == 0 ? mbits.x : (mw == 1 ? mbits.y : (mw == 2 ? mbits.z : mbits.w))) >> (ch0 & 31);
#pragma unroll
                  for (int i = 0; i < 16; ++i) v[i] = ((m16 >> i) & 1u) ? v[i] * L.mask_scale : 0.f;
                }
                if (L.out_cm != nullptr) {
#pragma unroll
                  for (int i = 0; i < 16; ++i)
                    if (ch0 + i < L.cout) L.out_cm[o0 + i * cs] = v[i];
                }
              }
              if (!last) {
                if (ch0 < L.next_k16) {
                  uint32_t H[8], Lo[8];
#pragma unroll
                  for (int i = 0; i < 8; ++i) split2(v[2 * i], v[2 * i + 1], H[i], Lo[i]);
                  const uint32_t base = out_buf + part_base<NT>(ch0 >> 6, 0) + row_base;
                  const int c8 = (ch0 & 63) >> 3;
                  const uint32_t a0 = base + (((c8) ^ r7) << 4), a1 = base + (((c8 + 1) ^ r7) << 4);
                  st_shared_v4(a0, H[0], H[1], H[2], H[3]);
                  st_shared_v4(a1, H[4], H[5], H[6], H[7]);
                  st_shared_v4(a0 + NT * 128, Lo[0], Lo[1], Lo[2], Lo[3]);
                  st_shared_v4(a1 + NT * 128, Lo[4], Lo[5], Lo[6], Lo[7]);
                }
              } else if (p.out_mode == CPFN_MLP_OUT_ROWS) {
                if (stage_rows) {
                  // a thread owns a ROW: writing it straight to global memory is 32 scattered 4-byte stores per warp
                  // instruction (the last, 35-channel layer of the heads chain cost more than a 128-channel one); the
                  // tile's rows go through the (now free) operand tile instead and leave as coalesced stores below
                  float4 *dst = reinterpret_cast<float4 *>(reinterpret_cast<float *>(act0) + row * row_ls + ch0);
#pragma unroll
                  for (int i4 = 0; i4 < 4; ++i4) dst[i4] = make_float4(v[i4 * 4], v[i4 * 4 + 1], v[i4 * 4 + 2], v[i4 * 4 + 3]);
                } else if (row_ok) {
                  float *o = p.out + col * p.ldo + ch0;
#pragma unroll
                  for (int i = 0; i < 16; ++i)
                    if (ch0 + i < L.cout) o[i] = v[i];
                }
              } else {
                // max over the warp's 32 rows: one REDUX per channel (post-ReLU floats order like uint32)
                uint32_t mine = 0u;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const uint32_t u = __reduce_max_sync(0xffffffffu, row_ok ? __float_as_uint(v[i]) : 0u);
                  if (lane == i) mine = u;
                }
                if (lane < 16 && ch0 + lane < L.cout && col0 + wq * 32 < p.cols)
                  atomicMax(reinterpret_cast<unsigned int *>(p.out) + ((col0 + wq * 32) / p.pool_g) * p.ldo + ch0 + lane, mine);
              }
            }
          }
          if (stage_rows && m0 + wave_max >= L.cout_chunks) {
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const long long left = p.cols - col0;
            const int nr = left < 128 ? static_cast<int>(left) : 128;
            const float *s_out = reinterpret_cast<const float *>(act0);
            float *g_out = p.out + col0 * p.ldo;
            const int t = w8 * 32 + lane, dr = 256 / L.cout, dc = 256 % L.cout;
            int r = t / L.cout, c = t % L.cout;
            while (r < nr) {
              g_out[static_cast<long long>(r) * p.ldo + c] = s_out[r * row_ls + c];
              r += dr; c += dc;
              if (c >= L.cout) { c -= L.cout; ++r; }
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");      // the next tile's loader overwrites the operand tile
          }
          tc_fence_before();
          const bool final_wave = last && (m0 + wave_max >= L.cout_chunks);
          if (!final_wave) {
            fence_proxy_async();
            mbar_arrive(act_ready);
          }
          pf.lap(p, 10 + min(l, 5));
        }
      }
    }
    if (pf.on) {
      const long long tiles = pf.acc[2];
      pf.flush(p, 5, 2, pf.now() - pf_begin);
      atomicAdd(p.prof + 8, 1ull);
      atomicAdd(p.prof + 9, static_cast<unsigned long long>(tiles));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

// Two sub-tiles of 128 columns per CTA: worker group g (4 warps, thread = point) owns sub-tile g from
// its gather to its last epilogue, the MMA thread alternates between the two, so the tensor pipe works
// on one sub-tile while the other is in its epilogue.
__global__ void __launch_bounds__(kChainThreads, 2)
mlp_chain_pm_kernel(const __grid_constant__ ChainP p) {
  constexpr int NT = 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t *ring = smem;
  uint8_t *act = ring + p.nstage * kStageBytes;                       // [2][act_bytes0]
  uint64_t *bars = reinterpret_cast<uint64_t *>(act + 2 * p.act_bytes0);
  uint64_t *full = bars, *empty = bars + kMaxStages, *act_ready = bars + 2 * kMaxStages,   // [2]
           *acc_full = bars + 2 * kMaxStages + 2;                                          // [2]
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * kMaxStages + 4);
  int *s_scratch = reinterpret_cast<int *>(bars + 32);                // per group: arow[128], brow[128][4], w[128][4]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.nstage; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    for (int g = 0; g < 2; ++g) { mbar_init(act_ready + g, 128); mbar_init(acc_full + g, 1); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // warp-uniform for the compiler
  const int sub_cols = p.tmem_cols / 2;                               // TMEM columns of one sub-tile
  const int n_pairs = (p.n_tiles + 1) / 2;

  if (warp == 0) {
    // ===== weight producer: layer by layer, once per valid sub-tile =====
    if (lane == 0) {
      Prof pf(p);
      const long long pf_begin = pf.now();
      int stage = 0;
      uint32_t phase = 0;
      for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
        const int nsub = (2 * pair + 1 < p.n_tiles) ? 2 : 1;
        int blk0 = 0;
        for (int l = 0; l < p.n_layers; ++l) {
          const int nblk = p.L[l].cout_chunks * p.L[l].cin_atoms * 2;
          for (int g = 0; g < nsub; ++g)
            for (int blk = 0; blk < nblk; ++blk) {
              pf.start();
              mbar_wait(empty + stage, phase ^ 1, p.wait_hint);
              pf.stop(0);
              mbar_arrive_expect_tx(full + stage, kStageBytes);
              bulk_g2s(ring + stage * kStageBytes, p.weights + static_cast<size_t>(blk0 + blk) * kStageBytes,
                       kStageBytes, full + stage);
              if (++stage == p.nstage) { stage = 0; phase ^= 1; }
            }
          blk0 += nblk;
        }
      }
      pf.flush(p, 3, 1, pf.now() - pf_begin);
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    {
      const uint32_t leader = elect_one();
      constexpr uint32_t idesc0 = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(128 >> 4) << 24);
      Prof pf(p);
      const long long pf_begin = pf.now();
      int stage = 0;
      uint32_t phase = 0, act_phase[2] = {0, 0};
      for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
        const int nsub = (2 * pair + 1 < p.n_tiles) ? 2 : 1;
        for (int l = 0; l < p.n_layers; ++l) {
          const LayerP &L = p.L[l];
          const int cout16 = (L.cout + 15) & ~15;
          for (int g = 0; g < nsub; ++g) {
            const uint32_t in_buf = smem_u32(act + g * p.act_bytes0);
            pf.start();
            mbar_wait(act_ready + g, act_phase[g], p.wait_hint);
            pf.stop(0);
            act_phase[g] ^= 1;
            tc_fence_after();
            for (int m = 0; m < L.cout_chunks; ++m) {
              const uint32_t d_tmem = tmem_base + g * sub_cols + m * 128;
              const int nch = min(128, cout16 - m * 128);
              const bool swap = p.pool_fast && l == p.n_layers - 1;        // pooled last layer: channels on the TMEM lanes
              const uint32_t idesc = idesc0 | (static_cast<uint32_t>((swap ? 128 : nch) >> 3) << 17);
              for (int j = 0; j < L.cin_atoms; ++j) {
                const int ks = min(4, L.ksteps - 4 * j);
                const uint32_t a_hi = in_buf + part_base<NT>(j, 0), a_lo = a_hi + NT * 128;
                pf.start();
                mbar_wait(full + stage, phase, p.wait_hint);            // W_hi block
                pf.stop(1);
                tc_fence_after();
                const uint64_t d_ahi = make_desc(a_hi), d_alo = make_desc(a_lo);
                uint64_t d_w = make_desc(smem_u32(ring + stage * kStageBytes));
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                  if (kk < ks) {
                    umma_bf16(d_tmem, swap ? d_w + 2 * kk : d_ahi + 2 * kk, swap ? d_ahi + 2 * kk : d_w + 2 * kk, idesc,
                              (j | kk) != 0 ? 1u : 0u, leader);
                    umma_bf16(d_tmem, swap ? d_w + 2 * kk : d_alo + 2 * kk, swap ? d_alo + 2 * kk : d_w + 2 * kk, idesc, 1u, leader);
                  }
                umma_commit(empty + stage, leader);
                if (++stage == p.nstage) { stage = 0; phase ^= 1; }
                pf.start();
                mbar_wait(full + stage, phase, p.wait_hint);            // W_lo block
                pf.stop(1);
                tc_fence_after();
                d_w = make_desc(smem_u32(ring + stage * kStageBytes));
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                  if (kk < ks)
                    umma_bf16(d_tmem, swap ? d_w + 2 * kk : d_ahi + 2 * kk, swap ? d_ahi + 2 * kk : d_w + 2 * kk, idesc, 1u, leader);
                umma_commit(empty + stage, leader);
                if (++stage == p.nstage) { stage = 0; phase ^= 1; }
              }
            }
            umma_commit(acc_full + g, leader);
          }
        }
      }
      if (leader) pf.flush(p, 0, 2, pf.now() - pf_begin);
    }
  } else {
    const int g = (warp - 2) >> 2;                    // worker group = sub-tile
    const int w4 = (warp - 2) & 3;
    const int wq = warp & 3;                          // TMEM lane quarter = rows wq*32 .. wq*32+31 of the sub-tile
    const int row = wq * 32 + lane;
    const uint32_t row_base = static_cast<uint32_t>((row >> 3) * 1024 + (row & 7) * 128);
    const int r7 = row & 7;
    const uint32_t my_act = smem_u32(act + g * p.act_bytes0);
    int *s_arow = s_scratch + g * 1152, *s_brow = s_arow + 128;
    float *s_w = reinterpret_cast<float *>(s_brow + 512);
    uint32_t acc_phase = 0;
    Prof pf(p);
    pf.on = pf.on && threadIdx.x == 64;
    const long long pf_begin = pf.now();
    for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
      const int tile = 2 * pair + g;
      if (tile >= p.n_tiles) continue;
      const long long col0 = tile_col0<NT>(p, tile);
      const long long col = col0 + row;
      const bool row_ok = col < p.cols;
      // thread = point: the cloud of THIS row (a tile may straddle two clouds when N is not a multiple of 128)
      const long long cloud = (row_ok ? col : col0) / p.cols_per_cloud;
      const long long n_in_cloud = (row_ok ? col : col0) - cloud * p.cols_per_cloud - row;
      pf.start();
      load_tile<NT, 4>(p, my_act, col0, w4, lane, s_arow, s_brow, s_w, 1 + g);
      fence_proxy_async();
      mbar_arrive(act_ready + g);
      pf.stop(0);
      pf.acc[2] += 1;
      for (int l = 0; l < p.n_layers; ++l) {
        const LayerP &L = p.L[l];
        const bool last = (l == p.n_layers - 1);
        const int cout_pad = L.cout_chunks * 128, cout16 = (L.cout + 15) & ~15;
        const bool slow = (L.mask_bits != nullptr) || (L.out_cm != nullptr);
        const float *bias_base = L.bias + (L.bias_per_cloud ? cloud * cout_pad : 0);
        uint4 mbits = make_uint4(~0u, ~0u, ~0u, ~0u);     // this point's keep bits (thread = point)
        if (L.mask_bits != nullptr && row_ok) {
          const uint32_t *mb = L.mask_bits + col * L.mask_words;
          mbits.x = __ldg(mb);
          if (L.mask_words > 1) mbits.y = __ldg(mb + 1);
          if (L.mask_words > 2) mbits.z = __ldg(mb + 2);
          if (L.mask_words > 3) mbits.w = __ldg(mb + 3);
        }
        pf.start();
        mbar_wait(acc_full + g, acc_phase, p.wait_hint);
        pf.stop(1);
        pf.start();
        acc_phase ^= 1;
        tc_fence_after();
        if (last && p.pool_fast) {
          // pooled last layer: channels on the TMEM lanes, points along the columns (see epi_pool_cols)
          const int n_groups = 128 / p.pool_g;
          for (int m = 0; m < L.cout_chunks; ++m) {
            const int ch = m * 128 + wq * 32 + lane;
            const bool ch_ok = ch < L.cout;
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + g * sub_cols + m * 128;
            for (int gi = 0; gi < n_groups; ++gi) {
              const long long gcol = col0 + static_cast<long long>(gi) * p.pool_g;
              const float bias = ch_ok ? __ldg(L.bias + (L.bias_per_cloud ? (gcol / p.cols_per_cloud) * cout_pad : 0) + ch) : 0.f;
              epi_pool_cols(taddr, gi * p.pool_g, p.pool_g, bias, L.relu != 0, p.out + (gcol / p.pool_g) * p.ldo + ch, ch_ok);
            }
          }
          tc_fence_before();
          pf.lap(p, 10 + min(l, 5));
          continue;      // last layer: nothing to publish (the next tile's loader meets this sub-tile's warps at its barrier)
        }
        if (!last && !slow) {           // the common layer: straight-line loop (see epi_next)
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + g * sub_cols;
          if (L.relu) epi_next<true>(taddr, bias_base, 0, 0, cout16, L.next_k16, my_act, row_base, r7);
          else epi_next<false>(taddr, bias_base, 0, 0, cout16, L.next_k16, my_act, row_base, r7);
          tc_fence_before();
          fence_proxy_async();
          mbar_arrive(act_ready + g);
          pf.lap(p, 10 + min(l, 5));
          continue;
        }
#pragma unroll 1
        for (int ch0 = 0; ch0 < cout16; ch0 += 16) {
          uint32_t r[16];
          float4 b4[4];                                // the bias loads are in flight while the accumulators arrive
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4) b4[i4] = __ldg(reinterpret_cast<const float4 *>(bias_base + ch0) + i4);
          tmem_ld16(tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + g * sub_cols + ch0, r);
          float v[16];
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4) {
            v[i4 * 4] = __uint_as_float(r[i4 * 4]) + b4[i4].x;
            v[i4 * 4 + 1] = __uint_as_float(r[i4 * 4 + 1]) + b4[i4].y;
            v[i4 * 4 + 2] = __uint_as_float(r[i4 * 4 + 2]) + b4[i4].z;
            v[i4 * 4 + 3] = __uint_as_float(r[i4 * 4 + 3]) + b4[i4].w;
          }
          if (L.relu) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
          }
          if (slow && row_ok) {                    // dropout mask / channel-major copy: coalesced over the warp
            const size_t o0 = (static_cast<size_t>(cloud) * L.cout + ch0) * p.cols_per_cloud + n_in_cloud + row;
            const size_t cs = static_cast<size_t>(p.cols_per_cloud);
            if (L.mask_bits != nullptr) {
              const int mw = ch0 >> 5;
              const uint32_t m16 = (mw == 0 ? mbits.x : (mw == 1 ? mbits.y : (mw == 2 ? mbits.z : mbits.w))) >> (ch0 & 31);
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = ((m16 >> i) & 1u) ? v[i] * L.mask_scale : 0.f;
            }
            if (L.out_cm != nullptr) {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (ch0 + i < L.cout) L.out_cm[o0 + i * cs] = v[i];
            }
          }
          if (!last) {
            if (ch0 < L.next_k16) {
              uint32_t H[8], Lo[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) split2(v[2 * i], v[2 * i + 1], H[i], Lo[i]);
              const uint32_t base = my_act + part_base<NT>(ch0 >> 6, 0) + row_base;
              const int c8 = (ch0 & 63) >> 3;
              const uint32_t a0 = base + (((c8) ^ r7) << 4), a1 = base + (((c8 + 1) ^ r7) << 4);
              st_shared_v4(a0, H[0], H[1], H[2], H[3]);
              st_shared_v4(a1, H[4], H[5], H[6], H[7]);
              st_shared_v4(a0 + NT * 128, Lo[0], Lo[1], Lo[2], Lo[3]);
              st_shared_v4(a1 + NT * 128, Lo[4], Lo[5], Lo[6], Lo[7]);
            }
          } else if (p.out_mode == CPFN_MLP_OUT_ROWS) {
            if (row_ok) {
              float *o = p.out + col * p.ldo + ch0;
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (ch0 + i < L.cout) o[i] = v[i];
            }
          } else {
            // max over the warp's 32 rows: one REDUX per channel (post-ReLU floats order like uint32)
            uint32_t mine = 0u;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const uint32_t u = __reduce_max_sync(0xffffffffu, row_ok ? __float_as_uint(v[i]) : 0u);
              if (lane == i) mine = u;
            }
            if (lane < 16 && ch0 + lane < L.cout && col0 + wq * 32 < p.cols)
              atomicMax(reinterpret_cast<unsigned int *>(p.out) + ((col0 + wq * 32) / p.pool_g) * p.ldo + ch0 + lane, mine);
          }
        }
        tc_fence_before();
        if (!last) {
          fence_proxy_async();
          mbar_arrive(act_ready + g);
        }
        pf.lap(p, 10 + min(l, 5));
      }
    }
    if (pf.on) {
      const long long tiles = pf.acc[2];
      pf.flush(p, 5, 2, pf.now() - pf_begin);
      atomicAdd(p.prof + 8, 1ull);
      atomicAdd(p.prof + 9, static_cast<unsigned long long>(tiles));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

uint32_t g_wait_hint = 0x989680u;
unsigned long long *g_prof_buf = nullptr;     // [kProfLaunches][kProfSlots], CPFN_CHAIN_PROFILE=1
int g_prof_launch = 0;

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
    L.cin_atoms = (s.cin + 63) / 64;
    L.ksteps = (s.cin + 15) / 16;
    L.cout_chunks = (s.cout + 127) / 128;
    L.cout = s.cout;
    L.relu = s.relu;
    L.bias_per_cloud = s.bias_per_cloud;
    L.bias = s.bias; L.out_cm = s.out_cm;
    L.mask_bits = s.mask_bits; L.mask_scale = s.mask_scale; L.mask_words = (s.cout + 31) / 32;
    if (s.mask_bits != nullptr && s.cout > 128) return CPFN_EINVAL;
    L.next_k16 = 0;
    if (l > 0) {
      if (s.cin != c->layers[l - 1].cout) return CPFN_EINVAL;
      p.L[l - 1].next_k16 = L.ksteps * 16;
    }
    total_blocks += L.cout_chunks * L.cin_atoms * 2;
    if (L.cout_chunks > max_chunks) max_chunks = L.cout_chunks;
    const size_t in_bytes = static_cast<size_t>(L.cin_atoms) * 2 * NT * 128;
    if (in_bytes > act_need[0]) act_need[0] = in_bytes;
    if (NT != 128 && (s.bias_per_cloud || s.mask_bits || s.out_cm) && (c->cols_per_cloud % NT) != 0) return CPFN_EINVAL;
  }
  if (static_cast<size_t>(total_blocks) * kStageBytes != c->weight_bytes) return CPFN_EINVAL;
  p.weights = static_cast<const uint8_t *>(c->weights);
  p.total_blocks = total_blocks;
  p.in_mode = c->in_mode; p.B = c->B; p.cols_per_cloud = c->cols_per_cloud;
  p.cols = static_cast<long long>(c->B) * c->cols_per_cloud;
  p.n_tiles = static_cast<int>((p.cols + NT - 1) / NT);
  p.win_cols = 0; p.win_off = 0;
  if (c->win_cols > 0) {
    // a window must be made of whole tiles and whole pooling groups, inside the cloud; pooled outputs are merged
    // with atomics into a buffer the CALLER zero-fills once (several windows of one output run as separate launches)
    if ((c->win_cols % NT) != 0 || (c->win_off % NT) != 0 || c->win_off < 0 || c->win_off + c->win_cols > c->cols_per_cloud ||
        (c->out_mode == CPFN_MLP_OUT_POOL && (!c->out_prezeroed || (c->win_cols % c->pool_g) != 0 || (c->win_off % c->pool_g) != 0)) ||
        c->split_cout)
      return CPFN_EINVAL;
    p.win_cols = c->win_cols; p.win_off = c->win_off;
    p.n_tiles = static_cast<int>(static_cast<long long>(c->B) * c->win_cols / NT);
  }
  p.a_src = c->a_src; p.a_ch = c->a_ch; p.a_rows = c->a_rows;
  p.idx = c->idx; p.xyz = c->xyz; p.centers = c->centers; p.group_k = c->group_k;
  p.b_src = c->b_src; p.b_ch = c->b_ch; p.b_rows = c->b_rows; p.nn_w = c->nn_w;
  p.out_mode = c->out_mode; p.out = c->out; p.ldo = c->ldo; p.pool_g = c->pool_g;
  // input row width must match layer 0; source row ids are kept as int32
  int width = c->a_ch;
  if (c->in_mode == CPFN_MLP_IN_GROUP) width += 3;
  else if (c->in_mode == CPFN_MLP_IN_INTERP) width += c->b_ch;
  p.l0_w = c->l0_w; p.l0_b = c->l0_b; p.l0_cout = c->l0_cout;
  p.in_bias = c->in_bias;
  if (c->in_bias != nullptr && c->in_mode != CPFN_MLP_IN_INTERP) return CPFN_EINVAL;
  p.xyz_w = c->xyz_w;
  if (c->xyz_w != nullptr) {
    if (NT != 128 || c->in_mode != CPFN_MLP_IN_GROUP || c->a_ch <= 0 || c->l0_w != nullptr || c->layers[0].bias_per_cloud ||
        c->layers[0].mask_bits || c->layers[0].out_cm || c->n_layers < 2)
      return CPFN_EINVAL;
    width = c->a_ch;
  }
  if (c->l0_w != nullptr) {
    if (c->in_mode != CPFN_MLP_IN_GROUP || c->a_ch != 0 || !c->l0_b || c->l0_cout <= 0 || (c->l0_cout & 3)) return CPFN_EINVAL;
    width = c->l0_cout;
  }
  if (width != c->layers[0].cin || (c->a_ch & 3) || (c->b_ch & 3)) return CPFN_EINVAL;
  if (c->a_ch > 0 && !c->a_src) return CPFN_EINVAL;
  if (c->in_mode == CPFN_MLP_IN_GROUP && (!c->idx || !c->xyz || !c->centers || c->group_k <= 0)) return CPFN_EINVAL;
  if (c->in_mode == CPFN_MLP_IN_INTERP && (!c->idx || !c->b_src || !c->nn_w)) return CPFN_EINVAL;
  if (p.cols >= 2147483647LL || static_cast<long long>(c->B) * c->a_rows >= 2147483647LL ||
      static_cast<long long>(c->B) * c->b_rows >= 2147483647LL) return CPFN_EINVAL;
  bool atomic_pool = false;
  p.pool_fast = 0;
  if (c->out_mode == CPFN_MLP_OUT_POOL && NT == 128) {
    // points-as-M kernel: every warp (32 rows) max-reduces and merges with atomicMax ...
    if (c->pool_g <= 0 || (c->pool_g % 32) != 0 || !c->layers[c->n_layers - 1].relu ||
        (c->cols_per_cloud % c->pool_g) != 0) return CPFN_EINVAL;
    atomic_pool = true;
    // ... unless whole pooling groups sit inside full tiles: then the last layer runs with the operand roles swapped
    // and every thread pools its own channel over the group's columns (epi_pool_cols: no atomics, no zero fill)
    const cpfn_mlp_layer_t &last = c->layers[c->n_layers - 1];
    const long long span = c->win_cols > 0 ? c->win_cols : c->cols_per_cloud;
    if ((c->pool_g == 32 || c->pool_g == 64 || c->pool_g == 128) && (span % 128) == 0 && !last.mask_bits && !last.out_cm)
      p.pool_fast = 1;
  } else if (c->out_mode == CPFN_MLP_OUT_POOL) {
    if (c->pool_g <= 0 || (c->pool_g % 16) != 0 || !c->layers[c->n_layers - 1].relu) return CPFN_EINVAL;
    if (c->pool_g <= NT / 2) {
      if ((NT / 2) % c->pool_g != 0) return CPFN_EINVAL;
    } else {
      if ((c->pool_g % (NT / 2)) != 0 || (c->cols_per_cloud % c->pool_g) != 0) return CPFN_EINVAL;
      atomic_pool = true;
    }
  }
  p.split_cout = (c->split_cout && NT != 128) ? 1 : 0;
  if (c->split_cout && (c->n_layers != 1 || NT == 128)) return CPFN_EINVAL;
  if (p.split_cout) max_chunks = 1;
  if (NT == 128 && max_chunks > 2) return CPFN_EINVAL;   // <= 256 TMEM columns per (sub-)tile
  // 128-column tiles: two sub-tiles per CTA (ping-pong) when that still lets two CTAs share an SM,
  // else one tile per CTA if THAT lets two CTAs share an SM, else two sub-tiles on one CTA per SM.
  const size_t max_smem = 227 * 1024;
  const size_t overhead = 1024 /*align*/ + kMiscBytes + 2 * kStageBytes;
  bool two_sub = true;
  if (NT == 128 && 2 * act_need[0] + overhead > max_smem / 2 - 1024 && act_need[0] + overhead <= max_smem / 2 - 1024 &&
      pow2_at_least(128 * max_chunks) <= 256)
    two_sub = false;
  if (NT == 128 && c->xyz_w != nullptr) two_sub = false;       // the position term lives in the one-tile kernel's epilogue
  p.tmem_cols = NT == 128 ? (two_sub ? 2 : 1) * pow2_at_least(128 * max_chunks)
                          : pow2_at_least(NT * (max_chunks < 512 / NT ? max_chunks : 512 / NT));
  p.act_bytes0 = static_cast<int>(act_need[0]);
  p.act_bytes1 = 0;
  // the epilogue writes a layer's output over its input, so every non-final layer must fit one TMEM wave
  for (int l = 0; l + 1 < c->n_layers; ++l)
    if (NT != 128 && p.L[l].cout_chunks * NT > p.tmem_cols) return CPFN_EINVAL;
  const size_t fixed = act_need[0] * ((NT == 128 && two_sub) ? 2 : 1) + 1024 /*align*/ + kMiscBytes;
  if (fixed + 2 * kStageBytes > max_smem) return CPFN_EINVAL;
  // Two CTAs per SM (one's epilogue overlaps the other's MMAs) when shared memory and TMEM allow.
  int per_sm = 1;
  int nstage = static_cast<int>((max_smem - fixed) / kStageBytes);
  const char *force1 = getenv("CPFN_CHAIN_ONE_PER_SM");      // tuning knob: deep weight ring instead of 2 CTAs/SM
  if (!(force1 && force1[0] == '1') && p.tmem_cols <= 256 && fixed + 2 * kStageBytes <= max_smem / 2 - 1024) {
    per_sm = 2;
    nstage = static_cast<int>((max_smem / 2 - 1024 - fixed) / kStageBytes);
  }
  if (nstage > kMaxStages) nstage = kMaxStages;
  if (nstage > total_blocks) nstage = total_blocks < 2 ? 2 : total_blocks;
  p.nstage = nstage;
  const size_t smem = fixed + static_cast<size_t>(nstage) * kStageBytes;
  void (*kern)(ChainP) = mlp_chain_kernel<NT>;
  if (NT == 128) kern = two_sub ? mlp_chain_pm_kernel : mlp_chain_pm1_kernel;
  if (NT != 128) p.pool_fast = 0;
  if (p.pool_fast && !two_sub && p.L[c->n_layers - 1].cout_chunks * 128 > p.tmem_cols) p.pool_fast = 0;   // one TMEM wave
  if (p.pool_fast) atomic_pool = false;
  CPFN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const int sms = sm_count() > 0 ? sm_count() : 148;
  const int units = (NT == 128 && two_sub) ? (p.n_tiles + 1) / 2 : p.n_tiles;   // the ping-pong kernel takes tile pairs
  int grid = units < per_sm * sms ? units : per_sm * sms;
  if (c->max_ctas > 0 && grid > c->max_ctas) grid = c->max_ctas;      // caller shares the GPU with another kernel
  if (grid <= 0) return CPFN_OK;
  if (atomic_pool && !c->out_prezeroed)
    CPFN_CUDA_TRY(cudaMemsetAsync(c->out, 0, sizeof(float) * static_cast<size_t>(p.cols / c->pool_g) * c->ldo, st));
  const dim3 grid2(grid, p.split_cout ? p.L[0].cout_chunks : 1);
  p.wait_hint = g_wait_hint;
  p.prof = nullptr;
  if (g_prof_buf != nullptr) {
    p.prof = g_prof_buf + static_cast<size_t>(g_prof_launch % kProfLaunches) * kProfSlots;
    ++g_prof_launch;
  }
  kern<<<grid2, kChainThreads, smem, st>>>(p);
  return check_launch();
}

}  // namespace
}  // namespace cpfn

// Host-side layout transform: W [cout, cin] row-major fp32 (BatchNorm already folded) ->
// blocks of 128 output channels x 64 input channels of bf16 in the kernel's shared-memory image
// (zero padded, 16-byte chunks XOR-swizzled by row & 7).  Every (chunk, K-atom) contributes a
// "hi" block (bf16(w)) followed by a "lo" block (bf16(w - hi)); blocks are ordered chunk-major
// then K-atom -- exactly the order the MMA thread consumes them.
static uint16_t cpfn_bf16_rn(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7F800000u) == 0x7F800000u) return static_cast<uint16_t>(u >> 16);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return static_cast<uint16_t>(u >> 16);
}

extern "C" size_t cpfn_mlp_packed_bytes(int cout, int cin) {
  if (cout <= 0 || cin <= 0) return 0;
  return static_cast<size_t>((cout + 127) / 128) * ((cin + 63) / 64) * 2 * cpfn::kStageBytes;
}

extern "C" int cpfn_mlp_pack_weights_host(const float *W, int cout, int cin, void *packed) {
  if (!W || !packed || cout <= 0 || cin <= 0) return CPFN_EINVAL;
  const int chunks = (cout + 127) / 128, atoms = (cin + 63) / 64;
  uint16_t *dst = static_cast<uint16_t *>(packed);
  const size_t blk_elems = cpfn::kStageBytes / 2;
  for (int m = 0; m < chunks; ++m)
    for (int j = 0; j < atoms; ++j) {
      uint16_t *hi = dst + (static_cast<size_t>(m) * atoms + j) * 2 * blk_elems;
      uint16_t *lo = hi + blk_elems;
      for (int r = 0; r < 128; ++r)
        for (int e = 0; e < 64; ++e) {
          const int co = m * 128 + r, ci = j * 64 + e;
          uint16_t h = 0, l = 0;
          if (co < cout && ci < cin) {
            const float f = W[static_cast<size_t>(co) * cin + ci];
            h = cpfn_bf16_rn(f);
            uint32_t hb = static_cast<uint32_t>(h) << 16;
            float hf;
            memcpy(&hf, &hb, 4);
            l = cpfn_bf16_rn(f - hf);
          }
          const int off = (r >> 3) * 1024 + (r & 7) * 128 + ((((e >> 3) ^ (r & 7)) << 4) | ((e & 7) << 1));
          hi[off >> 1] = h;
          lo[off >> 1] = l;
        }
    }
  return CPFN_OK;
}

extern "C" int cpfn_debug_chain_profile(unsigned long long *out, int max_launches, int reset) {
  using namespace cpfn;
  if (g_prof_buf == nullptr) return CPFN_EINVAL;
  CPFN_CUDA_TRY(cudaDeviceSynchronize());
  int n = g_prof_launch < kProfLaunches ? g_prof_launch : kProfLaunches;
  if (n > max_launches) n = max_launches;
  if (out != nullptr && n > 0)
    CPFN_CUDA_TRY(cudaMemcpy(out, g_prof_buf, sizeof(unsigned long long) * n * kProfSlots, cudaMemcpyDeviceToHost));
  if (reset) {
    CPFN_CUDA_TRY(cudaMemset(g_prof_buf, 0, sizeof(unsigned long long) * kProfLaunches * kProfSlots));
    g_prof_launch = 0;
  }
  return n;
}

extern "C" int cpfn_mlp_chain(const cpfn_mlp_chain_t *c, cpfn_stream_t stream) {
  using namespace cpfn;
  static bool hint_set = false;
  if (!hint_set) {
    hint_set = true;
    if (const char *e = getenv("CPFN_CHAIN_PROFILE")) {
      if (e[0] == '1' && cudaMalloc(&g_prof_buf, sizeof(unsigned long long) * kProfLaunches * kProfSlots) == cudaSuccess)
        cudaMemset(g_prof_buf, 0, sizeof(unsigned long long) * kProfLaunches * kProfSlots);
      else
        g_prof_buf = nullptr;
    }
    if (const char *e = getenv("CPFN_CHAIN_WAIT_HINT")) {
      g_wait_hint = static_cast<uint32_t>(strtoul(e, nullptr, 0));
    }
  }
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
