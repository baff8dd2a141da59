// Patch -> object merging: the device side of Utils/merging_utils.py and of the fusion
// block of evaluation_localSPFN.py:99-130.
//
// The reference materialises the dense point-to-primitive matrix A [N_global, M], M = nb*Kl + Kg
// (one column block per patch holding the patch's memberships on the rows of its points, one block with the
// object-level labels), and multiplies: similarity = A^T A (merging_utils.py:6-15, 128 GFLOP at 131 072
// points), fused labels = A @ one_hot(labels) (:46-50).  A is block-sparse: a point lies in ~2 of the 32
// patches.  Nothing dense is built here.  An inverse index  inv[b, p] = position of global point p in
// patch b, or -1  turns every product into gathers:
//   * A^T A block (b, b')  = sum over the points shared by patches b and b' of  W_b[j,:]^T W_b'[j',:]
//     (a [Kl x n_shared] x [n_shared x Kl] product; fp64 accumulators, one thread per entry);
//   * block (b, global)    = sum over the points of patch b of  W_b[j,:]^T S[p,:];  (global, global) = S^T S;
//   * fused labels, merged normals and types: one pass over the global points, looking the point up in
//     every patch in ascending patch order -- the order scatter_add_ visits them in the reference, so the
//     normal / type sums are bit-identical to the reference on the CPU.
#include <string.h>

#include <algorithm>
#include <utility>
#include <vector>

#include "common.cuh"

namespace cpfn {
namespace {

constexpr int kGramChunk = 32;       // hit rows staged per round (8 k-steps of the m8n8k4 DMMA)
constexpr int kMaxK = 32;            // label slots per patch / per object supported by the Gram kernel

__global__ void merge_inverse_kernel(const int32_t *__restrict__ patch_idx, int nb, int Np, int Ng,
                                     int32_t *__restrict__ inv) {
  const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= static_cast<long long>(nb) * Np) return;
  const int b = static_cast<int>(e / Np);
  const int p = __ldg(patch_idx + e);
  if (p >= 0 && p < Ng) inv[static_cast<size_t>(b) * Ng + p] = static_cast<int32_t>(e - static_cast<long long>(b) * Np);
}

constexpr int kGramThreads = 128;    // four warps: warp w owns the 8-row tile w of the block, all column tiles
constexpr int kGramHits = 1024;      // rows scanned per compaction round
static_assert(kGramThreads == 4 * kGramChunk, "four staging threads per row");
constexpr int kGramStride = 36;      // doubles per staged row: >= 32 and == 4 (mod 16), conflict-free fragment loads

// D (8x8, f64) += A (8x4, row) * B (4x8, col) on the FP64 tensor cores.  Fragments (PTX ISA, mma.m8n8k4 .f64):
// lane holds A[lane >> 2][lane & 3], B[lane & 3][lane >> 2], C[lane >> 2][2 * (lane & 3) + {0, 1}].
__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// One Gram block  G = sum_h rowA_h^T rowB_h  (KA x KB, KA, KB <= 32).
// kind 0: off-diagonal patch pairs (few shared rows: coarse row split); kind 1: diagonal blocks and (patch b,
// object labels) -- every row contributes, fine row split; kind 2: (object labels, object labels) over all
// points.  Every job's rows of the left operand are split over `parts` CTAs.  Per round: (1) the rows of the slice that have a
// partner in the right operand are compacted into a hit list (patches overlap in a few hundred points, not
// in 8192); (2) 32 hits at a time, both rows are staged in shared memory as fp64; (3) the warps run
// m8n8k4 DMMAs over them (A fragment = staged rows of the left operand read transposed).
__global__ void __launch_bounds__(kGramThreads)
merge_gram_kernel(const float *__restrict__ W, const int32_t *__restrict__ patch_idx, const float *__restrict__ S,
                  const int32_t *__restrict__ inv, int nb, int Np, int Kl, int Ng, int Kg, int split_dense,
                  int split_object, int split_sparse, double *__restrict__ acc) {
  __shared__ double sA[kGramChunk * kGramStride], sB[kGramChunk * kGramStride];
  __shared__ int s_ja[kGramHits], s_jb[kGramHits];
  __shared__ int s_n;
  const int M = nb * Kl + Kg;
  // One launch, three kinds of CTAs, long-running ones first: [dense | object-object | sparse pairs].
  int lin = blockIdx.x, kind, job, part, parts;
  if (lin < 2 * nb * split_dense) { kind = 1; parts = split_dense; job = lin / parts; part = lin - job * parts; }
  else if ((lin -= 2 * nb * split_dense) < split_object) { kind = 2; parts = split_object; job = 0; part = lin; }
  else { lin -= split_object; kind = 0; parts = split_sparse; job = lin / parts; part = lin - job * parts; }
  int b = 0, b2 = 0, mode = 0;                       // mode 0 patch-patch, 1 patch-object, 2 object-object
  if (kind == 0) {                                   // off-diagonal patch pairs b < b', row-major
    int r = job;
    while (r >= nb - 1 - b) { r -= nb - 1 - b; ++b; }
    b2 = b + 1 + r;
  } else if (kind == 1) {                            // every row has a partner: diagonal blocks, then patch-object
    if (job < nb) { b = b2 = job; } else { mode = 1; b = job - nb; }
  } else {
    mode = 2;
  }
  const int KA = mode == 2 ? Kg : Kl, KB = mode == 0 ? Kl : Kg;
  const float *A = mode == 2 ? S : W + static_cast<size_t>(b) * Np * Kl;
  const float *Bm = mode == 0 ? W + static_cast<size_t>(b2) * Np * Kl : S;
  const int32_t *pa = mode == 2 ? nullptr : patch_idx + static_cast<size_t>(b) * Np;
  const int32_t *invb = mode == 0 && b2 != b ? inv + static_cast<size_t>(b2) * Ng : nullptr;
  const int n_rows = mode == 2 ? Ng : Np;
  const int per = (n_rows + parts - 1) / parts;
  const int begin = part * per, end = min(n_rows, begin + per);
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int tiles_b = (KB + 7) >> 3;
  const bool computes = warp * 8 < KA;               // this warp's row tile exists
  double c[4][2];
#pragma unroll
  for (int q = 0; q < 4; ++q) c[q][0] = c[q][1] = 0.0;
  bool padded = false;                               // padding columns are zeroed before the first staging round
  for (int s0 = begin; s0 < end; s0 += kGramHits) {
    if (t == 0) s_n = 0;
    __syncthreads();
    const int s1 = min(end, s0 + kGramHits);
    if (mode == 0 && invb) {
      constexpr int kPer = kGramHits / kGramThreads;            // all index loads, then all look-ups, in flight together
      int pj[kPer], jb[kPer];
#pragma unroll
      for (int u = 0; u < kPer; ++u) { const int j = s0 + t + u * kGramThreads; pj[u] = j < s1 ? __ldg(pa + j) : -1; }
#pragma unroll
      for (int u = 0; u < kPer; ++u) jb[u] = pj[u] >= 0 ? __ldg(invb + pj[u]) : -1;   // position of the point in patch b'
#pragma unroll
      for (int u = 0; u < kPer; ++u)
        if (jb[u] >= 0) {
          const int slot = atomicAdd(&s_n, 1);
          s_ja[slot] = s0 + t + u * kGramThreads; s_jb[slot] = jb[u];
        }
    } else {                                                    // every row has its partner
      for (int j = s0 + t; j < s1; j += kGramThreads) {
        s_ja[j - s0] = j;
        s_jb[j - s0] = mode == 1 ? __ldg(pa + j) : j;           // the point's row of S | the row itself
      }
      if (t == 0) s_n = s1 - s0;
    }
    __syncthreads();
    const int n_all = s_n;
    if (n_all > 0 && !padded) {                                 // (uniform over the CTA)
      for (int i = t; i < kGramChunk * kGramStride; i += kGramThreads) sA[i] = sB[i] = 0.0;
      padded = true;
      __syncthreads();
    }
    for (int h0 = 0; h0 < n_all; h0 += kGramChunk) {
      const int n = min(kGramChunk, n_all - h0);
      const int n4 = (n + 3) & ~3;                              // rows up to the next k-step are zero-filled
      {                                                         // four threads per staged row, no divisions
        const int h = t >> 2, part = t & 3;
        if (h < n4) {
          const bool live = h < n;
          const float *ra = A + static_cast<size_t>(live ? s_ja[h0 + h] : 0) * KA;
          const float *rb = Bm + static_cast<size_t>(live ? s_jb[h0 + h] : 0) * KB;
          for (int col = part; col < KA; col += 4) sA[h * kGramStride + col] = live ? static_cast<double>(__ldg(ra + col)) : 0.0;
          for (int col = part; col < KB; col += 4) sB[h * kGramStride + col] = live ? static_cast<double>(__ldg(rb + col)) : 0.0;
        }
      }
      __syncthreads();
      if (computes) {
        for (int k0 = 0; k0 < n4; k0 += 4) {
          const int row = (k0 + (lane & 3)) * kGramStride + (lane >> 2);
          const double a = sA[row + warp * 8];
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (q < tiles_b) dmma_m8n8k4(c[q][0], c[q][1], a, sB[row + q * 8]);
        }
      }
      __syncthreads();
    }
  }
  if (computes) {
    const int row0 = mode == 2 ? nb * Kl : b * Kl;
    const int col0 = mode == 0 ? b2 * Kl : nb * Kl;
    const int ea = warp * 8 + (lane >> 2);
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int eb = q * 8 + 2 * (lane & 3) + i;
        if (ea < KA && eb < KB && c[q][i] != 0.0) atomicAdd(acc + static_cast<size_t>(row0 + ea) * M + col0 + eb, c[q][i]);
      }
  }
}

// fp64 accumulators (upper block triangle) -> symmetric fp32 matrix.
__global__ void merge_gram_finish_kernel(const double *__restrict__ acc, int nb, int Kl, int M, float *__restrict__ out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= M * M) return;
  const int i = e / M, j = e - i * M;
  const int bi = min(i / Kl, nb), bj = min(j / Kl, nb);            // column block of the entry (nb = object block)
  out[e] = static_cast<float>(bi <= bj ? acc[e] : acc[static_cast<size_t>(j) * M + i]);
}

// fused labels: out[p, l] = sum_c A[p,c] * v[label[c]] without A (evaluation_localSPFN.py:103-111 +
// merging_utils.get_point_final :46-50).  One warp per global point; its row of L sums lives in shared memory.
__global__ void __launch_bounds__(256)
merge_labels_kernel(const float *__restrict__ W, const float *__restrict__ S, const int32_t *__restrict__ inv,
                    const int32_t *__restrict__ labels, const float *__restrict__ v, int nb, int Np, int Kl, int Ng,
                    int Kg, int L, float *__restrict__ out) {
  extern __shared__ float s_acc[];                   // [warps][L]
  const int warp = threadIdx.x >> 5, lane = lane_id(), warps = blockDim.x >> 5;
  float *acc = s_acc + warp * L;
  for (int p = blockIdx.x * warps + warp; p < Ng; p += gridDim.x * warps) {
    for (int l = lane; l < L; l += 32) acc[l] = 0.f;
    __syncwarp();
    float cover = 0.f;                               // sum of the patch part of the row (:109)
    for (int b0 = 0; b0 < nb; b0 += 32) {
      const int bl = b0 + lane;
      const int jl = bl < nb ? __ldg(inv + static_cast<size_t>(bl) * Ng + p) : -1;
      unsigned int hits = __ballot_sync(0xffffffffu, jl >= 0);
      while (hits) {                                 // ascending patch order
        const int src = __ffs(hits) - 1;
        hits &= hits - 1;
        const int b = b0 + src;
        const int j = __shfl_sync(0xffffffffu, jl, src);
        const float *row = W + (static_cast<size_t>(b) * Np + j) * Kl;
        for (int k0 = 0; k0 < Kl; k0 += 32) {
          const int k = k0 + lane;
          const float w = k < Kl ? __ldg(row + k) : 0.f;
          // lanes of one patch normally carry distinct labels; lanes that do share one add in lane order
          const int lab = k < Kl ? __ldg(labels + b * Kl + k) : -1 - lane;
          const float wv = k < Kl ? __fmul_rn(w, __ldg(v + lab)) : 0.f;
          const unsigned int peers = __match_any_sync(0xffffffffu, lab);
          const int rank = __popc(peers & ((1u << lane) - 1u));
          const int rounds = __reduce_max_sync(0xffffffffu, k < Kl ? __popc(peers) : 0);
          for (int r = 0; r < rounds; ++r) {
            if (k < Kl && rank == r) acc[lab] = __fadd_rn(acc[lab], wv);
            __syncwarp();
          }
          float s = w;
          for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
          cover += s;
        }
      }
    }
    if (!(cover > 0.f)) {                            // not inside any patch: keep the object-level labels (:110)
      for (int g0 = 0; g0 < Kg; g0 += 32) {
        const int g = g0 + lane;
        const int lab = g < Kg ? __ldg(labels + nb * Kl + g) : -1 - lane;
        const float sv = g < Kg ? __fmul_rn(__ldg(S + static_cast<size_t>(p) * Kg + g), __ldg(v + lab)) : 0.f;
        const unsigned int peers = __match_any_sync(0xffffffffu, lab);
        const int rank = __popc(peers & ((1u << lane) - 1u));
        const int rounds = __reduce_max_sync(0xffffffffu, g < Kg ? __popc(peers) : 0);
        for (int r = 0; r < rounds; ++r) {
          if (g < Kg && rank == r) acc[lab] = __fadd_rn(acc[lab], sv);
          __syncwarp();
        }
      }
    }
    __syncwarp();
    for (int l = lane; l < L; l += 32) out[static_cast<size_t>(p) * L + l] = acc[l];
    __syncwarp();
  }
}

// get_point_final on a DENSE A (merging_utils.py:46-50): out[p,l] = sum over the columns c with label l of
// A[p,c] * v[l], columns in ascending order.  One warp per row; the row's L sums live in shared memory.
__global__ void __launch_bounds__(256)
merge_dense_labels_kernel(const float *__restrict__ A, const int32_t *__restrict__ labels, const float *__restrict__ v,
                          int Ng, int M, int L, float *__restrict__ out) {
  extern __shared__ float s_acc[];
  const int warp = threadIdx.x >> 5, lane = lane_id(), warps = blockDim.x >> 5;
  float *acc = s_acc + warp * L;
  for (int p = blockIdx.x * warps + warp; p < Ng; p += gridDim.x * warps) {
    for (int l = lane; l < L; l += 32) acc[l] = 0.f;
    __syncwarp();
    const float *row = A + static_cast<size_t>(p) * M;
    for (int c0 = 0; c0 < M; c0 += 32) {
      const int c = c0 + lane;
      const float a = c < M ? __ldg(row + c) : 0.f;
      const int lab = c < M ? __ldg(labels + c) : -1;
      const bool live = c < M && a != 0.f;                          // zero entries change nothing
      const int key = live ? lab : -1 - lane;
      const float av = live ? __fmul_rn(a, __ldg(v + lab)) : 0.f;
      const unsigned int peers = __match_any_sync(0xffffffffu, key);
      const int rank = __popc(peers & ((1u << lane) - 1u));
      const int rounds = __reduce_max_sync(0xffffffffu, live ? __popc(peers) : 0);
      for (int r = 0; r < rounds; ++r) {                            // columns sharing a label add in ascending order
        if (live && rank == r) acc[lab] = __fadd_rn(acc[lab], av);
        __syncwarp();
      }
    }
    for (int l = lane; l < L; l += 32) out[static_cast<size_t>(p) * L + l] = acc[l];
    __syncwarp();
  }
}

// Merged normals and types (evaluation_localSPFN.py:113-130): sums over the patches that contain the point, in
// ascending patch order; points no patch covers (all-zero normal sum) take the object-level prediction.
__global__ void __launch_bounds__(256)
merge_normals_types_kernel(const float *__restrict__ X, const float *__restrict__ T, const int32_t *__restrict__ inv,
                           const float *__restrict__ obj_normals, const float *__restrict__ obj_types, int nb, int Np,
                           int Ng, int n_types, float *__restrict__ out_normals, float *__restrict__ out_types) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= Ng) return;
  float nx = 0.f, ny = 0.f, nz = 0.f, den = 0.f;
  float num[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) num[t] = 0.f;
  for (int b = 0; b < nb; ++b) {
    const int j = __ldg(inv + static_cast<size_t>(b) * Ng + p);
    if (j < 0) continue;
    const size_t r = static_cast<size_t>(b) * Np + j;
    nx = __fadd_rn(nx, __ldg(X + r * 3)); ny = __fadd_rn(ny, __ldg(X + r * 3 + 1)); nz = __fadd_rn(nz, __ldg(X + r * 3 + 2));
#pragma unroll
    for (int t = 0; t < 8; ++t)
      if (t < n_types) num[t] = __fadd_rn(num[t], __ldg(T + r * n_types + t));
    den = __fadd_rn(den, 1.f);
  }
  const bool empty = nx == 0.f && ny == 0.f && nz == 0.f;          // torch.all(X_global == 0, axis=1)  (:117)
  if (empty) { nx = __ldg(obj_normals + static_cast<size_t>(p) * 3); ny = __ldg(obj_normals + static_cast<size_t>(p) * 3 + 1); nz = __ldg(obj_normals + static_cast<size_t>(p) * 3 + 2); }
  // F.normalize(p=2, dim=1, eps=1e-12): x / max(||x||, eps), the norm summed x^2 + y^2 + z^2 in that order
  const float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(nx, nx), __fmul_rn(ny, ny)), __fmul_rn(nz, nz)));
  const float dn = fmaxf(nrm, 1e-12f);
  out_normals[static_cast<size_t>(p) * 3] = __fdiv_rn(nx, dn);
  out_normals[static_cast<size_t>(p) * 3 + 1] = __fdiv_rn(ny, dn);
  out_normals[static_cast<size_t>(p) * 3 + 2] = __fdiv_rn(nz, dn);
  const float dd = fmaxf(den, 1.f);                                 // den.clamp(min=1)  (:127)
#pragma unroll
  for (int t = 0; t < 8; ++t)
    if (t < n_types)
      out_types[static_cast<size_t>(p) * n_types + t] = empty ? __ldg(obj_types + static_cast<size_t>(p) * n_types + t) : __fdiv_rn(num[t], dd);
}

}  // namespace
}  // namespace cpfn

using namespace cpfn;

extern "C" size_t cpfn_merge_inverse_bytes(int nb, int Ng) {
  return nb > 0 && Ng > 0 ? sizeof(int32_t) * static_cast<size_t>(nb) * Ng : 0;
}

extern "C" int cpfn_merge_inverse_index(const int32_t *patch_idx, int nb, int Np, int Ng, int32_t *inv,
                                        cpfn_stream_t stream) {
  if (!patch_idx || !inv || nb <= 0 || Np <= 0 || Ng <= 0) return CPFN_EINVAL;
  cudaStream_t s = as_stream(stream);
  CPFN_CUDA_TRY(cudaMemsetAsync(inv, 0xff, cpfn_merge_inverse_bytes(nb, Ng), s));
  const long long total = static_cast<long long>(nb) * Np;
  merge_inverse_kernel<<<static_cast<unsigned int>((total + 255) / 256), 256, 0, s>>>(patch_idx, nb, Np, Ng, inv);
  return check_launch();
}

extern "C" size_t cpfn_merge_similarity_workspace_bytes(int nb, int Kl, int Kg) {
  const size_t M = static_cast<size_t>(nb) * Kl + Kg;
  return sizeof(double) * M * M;
}

extern "C" int cpfn_merge_similarity(const float *W, const int32_t *patch_idx, const float *S, const int32_t *inv,
                                     int nb, int Np, int Kl, int Ng, int Kg, float *out, void *workspace,
                                     size_t workspace_bytes, cpfn_stream_t stream) {
  if (!W || !patch_idx || !S || !inv || !out || nb <= 0 || Np <= 0 || Ng <= 0 || Kl <= 0 || Kg <= 0 || Kl > kMaxK ||
      Kg > kMaxK)
    return CPFN_EINVAL;
  const size_t need = cpfn_merge_similarity_workspace_bytes(nb, Kl, Kg);
  if (!workspace || workspace_bytes < need || (reinterpret_cast<uintptr_t>(workspace) & 7)) return CPFN_EWORKSPACE;
  cudaStream_t s = as_stream(stream);
  double *acc = static_cast<double *>(workspace);
  CPFN_CUDA_TRY(cudaMemsetAsync(acc, 0, need, s));
  const int M = nb * Kl + Kg;
  const int sms = sm_count();
  if (sms <= 0) return CPFN_ELAUNCH;
  const int split_sparse = (Np + kGramHits - 1) / kGramHits;       // b < b': scan kGramHits rows per CTA
  const int split_dense = (Np + 8 * kGramChunk - 1) / (8 * kGramChunk);   // dense blocks: eight staging rounds per CTA
  int split_object = 8 * sms;                                      // the object-object block runs over all points
  const int gmax = (Ng + 8 * kGramChunk - 1) / (8 * kGramChunk);
  if (split_object > gmax) split_object = gmax;
  const long long ctas = 2ll * nb * split_dense + split_object + 1ll * (nb * (nb - 1) / 2) * split_sparse;
  if (ctas > 0x7fffffffll) return CPFN_EINVAL;
  merge_gram_kernel<<<static_cast<unsigned int>(ctas), kGramThreads, 0, s>>>(W, patch_idx, S, inv, nb, Np, Kl, Ng, Kg,
                                                                            split_dense, split_object, split_sparse, acc);
  merge_gram_finish_kernel<<<(M * M + 255) / 256, 256, 0, s>>>(acc, nb, Kl, M, out);
  return check_launch();
}

extern "C" int cpfn_merge_point_labels(const float *W, const float *S, const int32_t *inv, const int32_t *labels,
                                       const float *label_weight, int nb, int Np, int Kl, int Ng, int Kg, int L,
                                       float *out, cpfn_stream_t stream) {
  if (!W || !S || !inv || !labels || !label_weight || !out || nb <= 0 || Np <= 0 || Ng <= 0 || Kl <= 0 || Kg <= 0 ||
      L <= 0 || L > 1536)
    return CPFN_EINVAL;
  const int sms = sm_count();
  if (sms <= 0) return CPFN_ELAUNCH;
  int grid = (Ng + 7) / 8;
  if (grid > 8 * sms) grid = 8 * sms;
  merge_labels_kernel<<<grid, 256, sizeof(float) * 8 * L, as_stream(stream)>>>(W, S, inv, labels, label_weight, nb, Np,
                                                                              Kl, Ng, Kg, L, out);
  return check_launch();
}

extern "C" int cpfn_merge_dense_labels(const float *A, const int32_t *labels, const float *label_weight, int Ng, int M,
                                       int L, float *out, cpfn_stream_t stream) {
  if (!A || !labels || !label_weight || !out || Ng <= 0 || M <= 0 || L <= 0 || L > 1536) return CPFN_EINVAL;
  const int sms = sm_count();
  if (sms <= 0) return CPFN_ELAUNCH;
  int grid = (Ng + 7) / 8;
  if (grid > 8 * sms) grid = 8 * sms;
  merge_dense_labels_kernel<<<grid, 256, sizeof(float) * 8 * L, as_stream(stream)>>>(A, labels, label_weight, Ng, M, L,
                                                                                    out);
  return check_launch();
}

extern "C" int cpfn_merge_normals_types(const float *X, const float *T, const int32_t *inv, const float *obj_normals,
                                        const float *obj_types, int nb, int Np, int Ng, int n_types, float *out_normals,
                                        float *out_types, cpfn_stream_t stream) {
  if (!X || !T || !inv || !obj_normals || !obj_types || !out_normals || !out_types || nb <= 0 || Np <= 0 || Ng <= 0 ||
      n_types <= 0 || n_types > 8)
    return CPFN_EINVAL;
  merge_normals_types_kernel<<<(Ng + 255) / 256, 256, 0, as_stream(stream)>>>(X, T, inv, obj_normals, obj_types, nb, Np,
                                                                             Ng, n_types, out_normals, out_types);
  return check_launch();
}

// ---- host side: the greedy merge of merging_utils.heuristic_merging (:17-33) -----------------------------
// The reference repeatedly takes the arg-max pair, merges the two segments, ORs their patch sets and drops
// every pair whose segments now share a patch.  Patch sets only grow, so a dropped pair stays dropped: the
// same result comes from ONE pass over the pairs in stable descending order of the penalty, merging a pair
// iff its segments' patch sets are disjoint when it is reached -- except the very first pair (the first
// maximum in the reference's pair order), which the reference merges before any filtering.  Pairs inside one
// patch can therefore only ever merge as that first pair and are not even sorted.
namespace {

struct PairRec {
  uint64_t key;       // order-preserving bits of the penalty, inverted: ascending keys == descending penalties
  int32_t a, b;
};

inline uint64_t descending_key(double v) {
  uint64_t k;
  memcpy(&k, &v, sizeof(k));
  k = (k >> 63) ? ~k : (k | (1ull << 63));
  return ~k;
}
inline uint64_t descending_key(float v) {           // 32 significant bits: three radix passes instead of six
  uint32_t k;
  memcpy(&k, &v, sizeof(k));
  k = (k >> 31) ? ~k : (k | (1u << 31));
  return static_cast<uint32_t>(~k);
}

// Stable LSD radix sort by key over `bits` key bits, 8 bits per pass (256 output streams stay in L1; 2048 of
// them did not and made the scatter the slowest part of the solve); passes whose digit is constant are skipped.
// Scratch vectors live per host thread: a solve needs ~2 MB of them, and fresh allocations of that size are
// mapped and page-faulted on every call (measured: most of the solve's time in a container).
std::vector<PairRec> &scratch(int which) {
  static thread_local std::vector<PairRec> v[2];
  return v[which];
}

void radix_sort_records(std::vector<PairRec> &rec, int bits) {
  std::vector<PairRec> &tmp = scratch(1);
  tmp.resize(rec.size());
  size_t count[257];
  for (int shift = 0; shift < bits; shift += 8) {
    std::fill(count, count + 257, 0);
    for (const PairRec &r : rec) ++count[((r.key >> shift) & 255) + 1];
    bool single = false;
    for (int d = 1; d <= 256 && !single; ++d) single = count[d] == rec.size();
    if (single) continue;
    for (int d = 1; d <= 256; ++d) count[d] += count[d - 1];
    for (const PairRec &r : rec) tmp[count[(r.key >> shift) & 255]++] = r;
    rec.swap(tmp);
  }
}

class Segments {                                       // segment labels + the patch set of every label
 public:
  Segments(const int64_t *patch_id, int64_t n_nodes, int64_t *segment_id) : n_(n_nodes), seg_(segment_id) {
    int64_t n_patch = 0;
    for (int64_t i = 0; i < n_; ++i) n_patch = std::max(n_patch, patch_id[i] + 1);
    words_ = static_cast<size_t>((n_patch + 63) / 64);
    mask_.assign(static_cast<size_t>(n_) * words_, 0);
    for (int64_t i = 0; i < n_; ++i) {
      seg_[i] = i;
      mask_[static_cast<size_t>(i) * words_ + static_cast<size_t>(patch_id[i] / 64)] |= 1ull << (patch_id[i] % 64);
    }
  }
  // merges the segment of b into the segment of a (as :24-27 do); `unconditional` skips the patch-set test
  void offer(int64_t a, int64_t b, bool unconditional) {
    const int64_t la = seg_[a], lb = seg_[b];
    uint64_t *ma = mask_.data() + static_cast<size_t>(la) * words_, *mb = mask_.data() + static_cast<size_t>(lb) * words_;
    if (!unconditional)
      for (size_t w = 0; w < words_; ++w)
        if (ma[w] & mb[w]) return;
    if (la == lb) return;
    for (int64_t i = 0; i < n_; ++i)
      if (seg_[i] == lb) seg_[i] = la;
    for (size_t w = 0; w < words_; ++w) ma[w] |= mb[w];
  }

 private:
  int64_t n_;
  int64_t *seg_;
  size_t words_ = 1;
  std::vector<uint64_t> mask_;
};

// `rec`: the cross-patch pairs in the reference's pair order; (fa, fb): the first maximum over ALL pairs when it
// is a pair inside one patch (else fa < 0: the first maximum is then simply the first record after the sort).
void greedy_merge(std::vector<PairRec> &rec, int key_bits, int64_t fa, int64_t fb, Segments &segs) {
  radix_sort_records(rec, key_bits);                   // stable: ties keep the reference's order
  bool first = true;
  if (fa >= 0) { segs.offer(fa, fb, true); first = false; }
  for (const PairRec &r : rec) {
    segs.offer(r.a, r.b, first);
    first = false;
  }
}

bool valid_patch_ids(const int64_t *patch_id, int64_t n_nodes) {
  for (int64_t i = 0; i < n_nodes; ++i)
    if (patch_id[i] < 0) return false;
  return true;
}

// The pair list of run_heuristic_solver (:37-39: similarity > threshold, i < j, np.where order) straight from
// the matrix, then the greedy merge.
template <typename T>
int solve_from_matrix(const T *similarity, int64_t n_nodes, T threshold, const int64_t *patch_id, int64_t *segment_id) {
  if (n_nodes <= 0 || n_nodes > 0x7fffffff || !similarity || !patch_id || !segment_id || !valid_patch_ids(patch_id, n_nodes))
    return CPFN_EINVAL;
  std::vector<PairRec> &rec = scratch(0);
  rec.clear();
  bool have_best = false, best_same_patch = false;
  T best = 0;
  int64_t fa = -1, fb = -1;
  for (int64_t i = 0; i < n_nodes; ++i) {
    const T *row = similarity + i * n_nodes;
    const int64_t pi = patch_id[i];
    for (int64_t j = i + 1; j < n_nodes; ++j) {
      const T v = row[j];
      if (!(v > threshold)) continue;
      const bool same = patch_id[j] == pi;
      if (!have_best || v > best) { have_best = true; best = v; best_same_patch = same; fa = i; fb = j; }
      if (!same) rec.push_back(PairRec{descending_key(v), static_cast<int32_t>(i), static_cast<int32_t>(j)});
    }
  }
  Segments segs(patch_id, n_nodes, segment_id);
  if (!best_same_patch) fa = fb = -1;
  greedy_merge(rec, sizeof(T) == 4 ? 32 : 64, fa, fb, segs);
  return CPFN_OK;
}

}  // namespace

extern "C" int cpfn_heuristic_merging_host(const int64_t *pairs, const double *penalty, int64_t n_pairs,
                                           const int64_t *patch_id, int64_t n_nodes, int64_t *segment_id) {
  if (n_nodes <= 0 || n_nodes > 0x7fffffff || !patch_id || !segment_id || n_pairs < 0 || (n_pairs > 0 && (!pairs || !penalty)) ||
      !valid_patch_ids(patch_id, n_nodes))
    return CPFN_EINVAL;
  std::vector<PairRec> &rec = scratch(0);
  rec.clear();
  int64_t first_pair = -1;
  for (int64_t i = 0; i < n_pairs; ++i) {
    const int64_t a = pairs[2 * i], b = pairs[2 * i + 1];
    if (a < 0 || b < 0 || a >= n_nodes || b >= n_nodes) return CPFN_EINVAL;
    if (first_pair < 0 || penalty[i] > penalty[first_pair]) first_pair = i;
  }
  int64_t fa = -1, fb = -1;
  for (int64_t i = 0; i < n_pairs; ++i) {
    const int64_t a = pairs[2 * i], b = pairs[2 * i + 1];
    if (patch_id[a] != patch_id[b]) rec.push_back(PairRec{descending_key(penalty[i]), static_cast<int32_t>(a), static_cast<int32_t>(b)});
    else if (i == first_pair) { fa = a; fb = b; }
  }
  Segments segs(patch_id, n_nodes, segment_id);
  greedy_merge(rec, 64, fa, fb, segs);
  return CPFN_OK;
}

extern "C" int cpfn_merge_solve_host(const double *similarity, int64_t n_nodes, double threshold,
                                     const int64_t *patch_id, int64_t *segment_id) {
  return solve_from_matrix<double>(similarity, n_nodes, threshold, patch_id, segment_id);
}

extern "C" int cpfn_merge_solve_host_f32(const float *similarity, int64_t n_nodes, float threshold,
                                         const int64_t *patch_id, int64_t *segment_id) {
  return solve_from_matrix<float>(similarity, n_nodes, threshold, patch_id, segment_id);
}
