// Patch -> object label merge on the device (the solver half).
//
// Replaces the host step of the fusion block: Utils/merging_utils.py:35-44 (run_heuristic_solver: np.where over
// the similarity matrix, pair list, heuristic_merging, label replacement, np.unique) and :17-33 (the numba greedy
// loop) of the reference, which evaluation_localSPFN.py:102 runs on the CPU after a device->host copy of the
// matrix.  Same result, label for label:
//   1. pair keys: every entry (i < j) of the similarity matrix with value > threshold whose nodes lie in different
//      patches becomes a 64-bit key [order-inverted bits of the value | i*M + j]; ascending keys = descending
//      values, ties in the reference's np.where (row-major) order.  Pairs inside one patch can only ever merge as
//      the very first arg-max of the reference's loop (before it filters anything), so they are not sorted: one
//      block-wide arg-max over ALL pairs finds that first pair.
//   2. the keys are sorted (radix sort, cub::DeviceRadixSort on the 51 significant bits).
//   3. one CTA walks the sorted pairs 1024 at a time.  The reference takes the arg-max pair, merges its two
//      segments, ORs their patch sets and drops every pair whose segments now share a patch; patch sets only grow,
//      so a dropped pair stays dropped and ONE pass in descending order that merges a pair iff its segments' patch
//      sets are disjoint when it is reached gives the same segments.  Within a chunk all pairs are tested in
//      parallel, the first mergeable one is merged (relabel + OR of the patch sets, all threads), the pairs behind
//      it are re-tested.
//   4. labels of empty slots (diagonal < threshold) are replaced as :41-43 do, labels are made consecutive in the
//      order of np.unique (sorted distinct values), and 1 / (members + 1e-10) per label is what get_point_final
//      (:46-50) divides by.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace cpfn {
namespace {

constexpr int kSolveThreads = 256;
constexpr int kMaxNodes = 4096;          // nodes of the merge graph (nb * Kl + Kg); 700 in the reference's configuration.  Node ids are packed
                                         // into 16 bits and pair ids (i*M + j) into 32 in the greedy pass.
constexpr unsigned long long kNoPair = ~0ull;

struct SolveHeader {
  unsigned long long first_key;          // smallest key over ALL pairs (same-patch pairs included)
  unsigned int n_pairs;                  // cross-patch pairs written to the key list
  unsigned int pad;
};

__device__ __forceinline__ unsigned int descending_bits(float v) {
  unsigned int k = __float_as_uint(v);
  k = (k >> 31) ? ~k : (k | 0x80000000u);
  return ~k;
}

__device__ __forceinline__ int patch_of(int node, int nb, int Kl) { return node < nb * Kl ? node / Kl : nb; }

// One thread per matrix entry of the strict upper triangle (row-major).  Valid cross-patch pairs are appended to
// `keys` (order irrelevant: the sort restores it, keys are unique); every CTA also folds its smallest key over all
// valid pairs into header->first_key.
__global__ void __launch_bounds__(256)
solve_pairs_kernel(const float *__restrict__ sim, int M, int nb, int Kl, float threshold, int pair_bits,
                   unsigned long long *__restrict__ keys, SolveHeader *__restrict__ hdr) {
  const long long e = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  unsigned long long key = kNoPair;
  bool cross = false;
  if (e < static_cast<long long>(M) * M) {
    const int i = static_cast<int>(e / M), j = static_cast<int>(e - static_cast<long long>(i) * M);
    if (j > i) {
      const float v = __ldg(sim + e);
      if (v > threshold) {
        key = (static_cast<unsigned long long>(descending_bits(v)) << pair_bits) | static_cast<unsigned long long>(e);
        cross = patch_of(i, nb, Kl) != patch_of(j, nb, Kl);
      }
    }
  }
  // block-wide minimum of the keys (64-bit: two REDUX-free shuffles per step)
  unsigned long long m = key;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, m, o);
    m = other < m ? other : m;
  }
  __shared__ unsigned long long s_min[8];
  __shared__ unsigned int s_base, s_cnt[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned int ballot = __ballot_sync(0xffffffffu, cross);
  if (lane == 0) { s_min[warp] = m; s_cnt[warp] = __popc(ballot); }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long mm = s_min[0];
    unsigned int total = 0;
    for (int w = 0; w < 8; ++w) {
      mm = s_min[w] < mm ? s_min[w] : mm;
      const unsigned int c = s_cnt[w];
      s_cnt[w] = total;
      total += c;
    }
    if (mm != kNoPair) atomicMin(&hdr->first_key, mm);
    s_base = total ? atomicAdd(&hdr->n_pairs, total) : 0u;
  }
  __syncthreads();
  if (cross) keys[s_base + s_cnt[warp] + __popc(ballot & ((1u << lane) - 1u))] = key;
}

__global__ void solve_init_kernel(SolveHeader *hdr) {
  hdr->first_key = kNoPair;
  hdr->n_pairs = 0;
  hdr->pad = 0;
}

// The greedy pass + label post-processing: one CTA.  The pass itself is sequential in the merges, so ONE warp walks
// the sorted pairs: every lane holds four pairs and tests them against the current segments (a pair whose patch sets
// intersect is dead for good), the first live pair is merged, the pairs behind it are re-tested.  No block-wide
// barrier sits inside the pass.  The other seven warps stream the next block of sorted keys from global memory into a
// shared-memory double buffer as packed (a, b) node pairs -- and drop the pairs that are already dead in whatever
// state they see (patch sets only grow, so a stale or half-updated view can only under-estimate them: dead there is
// dead for good), which leaves warp 0 a small fraction of the late blocks.
constexpr int kBlockPairs = 4096;
constexpr int kLoaders = kSolveThreads / 32 - 1;                                   // 7 staging warps
constexpr int kSegPairs = ((kBlockPairs + kLoaders - 1) / kLoaders + 31) / 32 * 32;   // pairs per loader warp and block

template <bool kRegLabels>      // M <= 1024: every lane also keeps the labels of its 32 nodes in registers
__global__ void __launch_bounds__(kSolveThreads)
solve_greedy_kernel(const float *__restrict__ sim, const unsigned long long *__restrict__ sorted,
                    const SolveHeader *__restrict__ hdr, int M, int nb, int Kl, int Kg, float threshold, int pair_bits,
                    int32_t *__restrict__ labels, float *__restrict__ weights, int32_t *__restrict__ n_labels,
                    int32_t *__restrict__ segments) {
  extern __shared__ unsigned char s_raw[];
  unsigned long long *s_mask = reinterpret_cast<unsigned long long *>(s_raw);        // [M] patch set of a label
  int *s_seg = reinterpret_cast<int *>(s_mask + M);                                  // [M] label of a node
  int *s_aux = s_seg + M;                                                            // [M + Kmax] presence / ranks / counts
  unsigned int *s_pairs = reinterpret_cast<unsigned int *>(s_aux + M + (Kl > Kg ? Kl : Kg));   // [2][kLoaders][kSegPairs]: a | b << 16
  __shared__ int s_count[2][kLoaders];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned long long pair_mask = (1ull << pair_bits) - 1ull;
  for (int i = tid; i < M; i += kSolveThreads) {
    s_seg[i] = i;
    s_mask[i] = 1ull << patch_of(i, nb, Kl);
  }
  const unsigned int n_pairs = hdr->n_pairs;
  const unsigned int n_blocks = (n_pairs + kLoaders * kSegPairs - 1) / (kLoaders * kSegPairs);
  __syncthreads();

  auto dead = [&](int a, int b) { return (s_mask[s_seg[a]] & s_mask[s_seg[b]]) != 0ull; };

  // loader warp w (1..7) stages its segment of block blk: keys -> (a, b), dead pairs dropped, order kept
  auto stage = [&](unsigned int blk) {
    const int w = warp - 1;
    unsigned int *dst = s_pairs + ((blk & 1u) * kLoaders + w) * kSegPairs;
    const unsigned int base = (blk * kLoaders + w) * kSegPairs;
    int kept = 0;
    for (int c = 0; c < kSegPairs; c += 32) {
      const unsigned int p = base + c + lane;
      const bool valid = p < n_pairs;
      unsigned int ab = 0u;
      bool keep = false;
      if (valid) {
        const unsigned int e = static_cast<unsigned int>(sorted[p] & pair_mask);
        const unsigned int a = e / static_cast<unsigned int>(M), b = e - a * static_cast<unsigned int>(M);
        ab = a | (b << 16);
        keep = !dead(static_cast<int>(a), static_cast<int>(b));
      }
      const unsigned int bal = __ballot_sync(0xffffffffu, keep);
      if (keep) dst[kept + __popc(bal & ((1u << lane) - 1u))] = ab;
      kept += __popc(bal);
    }
    if (lane == 0) s_count[blk & 1u][w] = kept;
  };

  // s_seg[i] = current label of node i (flat: a test is two independent two-load chains per pair, which the four
  // pairs a lane holds overlap).  A merge relabels b's segment to a's label (merging_utils.py:24-27).  Warp 0 only.
  int reg_label[32];                                       // labels of nodes u * 32 + lane (kRegLabels)
#pragma unroll
  for (int u = 0; u < 32; ++u) reg_label[u] = u * 32 + lane;
  auto merge = [&](int a, int b) {                         // all lanes pass the same (a, b)
    const int la = s_seg[a], lb = s_seg[b];
    __syncwarp();
    if (la != lb) {
      if (kRegLabels) {                                    // compare in registers, store only what changes
#pragma unroll
        for (int u = 0; u < 32; ++u)
          if (u * 32 < M && reg_label[u] == lb) { reg_label[u] = la; s_seg[u * 32 + lane] = la; }
      } else {
        for (int i0 = 0; i0 < M; i0 += 256) {
          int v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int i = i0 + u * 32 + lane;
            v[u] = i < M ? s_seg[i] : -1;
          }
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (v[u] == lb) s_seg[i0 + u * 32 + lane] = la;
        }
      }
      if (lane == 0) s_mask[la] |= s_mask[lb];
    }
    __syncwarp();
  };

  if (warp == 0) {
    // the reference's first arg-max is merged before anything is filtered: if it is a pair inside one patch it can
    // only merge here (a cross-patch first pair is simply the first sorted key and passes the test anyway)
    const unsigned long long fk = hdr->first_key;
    if (fk != kNoPair) {
      const unsigned int e = static_cast<unsigned int>(fk & pair_mask);
      const int a = static_cast<int>(e / static_cast<unsigned int>(M)), b = static_cast<int>(e) - a * M;
      if (patch_of(a, nb, Kl) == patch_of(b, nb, Kl)) merge(a, b);
    }
  }
  __syncthreads();
  if (warp != 0 && n_blocks > 0) stage(0);
  __syncthreads();
  constexpr int kSub = 4;                                  // 4 x 32 pairs per step: the smem chains of the tests overlap
  for (unsigned int blk = 0; blk < n_blocks; ++blk) {
    if (warp == 0) {
      for (int w = 0; w < kLoaders; ++w) {
        const unsigned int *src = s_pairs + ((blk & 1u) * kLoaders + w) * kSegPairs;
        const int count = s_count[blk & 1u][w];
        for (int c0 = 0; c0 < count; c0 += 32 * kSub) {
          int a[kSub], b[kSub];
          bool alive[kSub];
#pragma unroll
          for (int k = 0; k < kSub; ++k) {
            const int p = c0 + k * 32 + lane;              // pair order: sub-chunk k, then lane
            alive[k] = p < count;
            const unsigned int ab = alive[k] ? src[p] : 0u;
            a[k] = static_cast<int>(ab & 0xFFFFu);
            b[k] = static_cast<int>(ab >> 16);
          }
          {                                                // a pair whose patch sets intersect is dead for good
            bool d[kSub];
#pragma unroll
            for (int k = 0; k < kSub; ++k) d[k] = dead(a[k], b[k]);   // unconditional: the load chains overlap
#pragma unroll
            for (int k = 0; k < kSub; ++k) alive[k] = alive[k] && !d[k];
          }
#pragma unroll
          for (int k = 0; k < kSub; ++k) {
            unsigned int behind = 0xFFFFFFFFu;             // lanes of sub-chunk k whose pair is still to be settled
            while (true) {
              const unsigned int live = __ballot_sync(0xffffffffu, alive[k]) & behind;
              if (live == 0u) break;
              const int pick = __ffs(live) - 1;
              merge(__shfl_sync(0xffffffffu, a[k], pick), __shfl_sync(0xffffffffu, b[k], pick));
              behind = pick == 31 ? 0u : (0xFFFFFFFFu << (pick + 1));
              // the state changed: re-test what is still ahead (a settled pair's flag is never read again)
              bool d[kSub];
#pragma unroll
              for (int q = 0; q < kSub; ++q) d[q] = q >= k ? dead(a[q], b[q]) : false;
#pragma unroll
              for (int q = 0; q < kSub; ++q) alive[q] = alive[q] && !d[q];
            }
          }
        }
      }
    } else if (blk + 1 < n_blocks) {
      stage(blk + 1);
    }
    __syncthreads();
  }

  // ---- labels: replacement of the empty slots (:41-43), np.unique(return_inverse) (:44), member counts ----
  const int Kmax = Kl > Kg ? Kl : Kg;
  for (int i = tid; i < M + Kmax; i += kSolveThreads) s_aux[i] = 0;
  __syncthreads();
  for (int i = tid; i < M; i += kSolveThreads) {
    int lab = s_seg[i];
    if (segments != nullptr) segments[i] = lab;
    if (__ldg(sim + static_cast<size_t>(i) * M + i) < threshold) lab = i < nb * Kl ? (i % Kl) - Kl : (i - nb * Kl) - Kg;
    s_seg[i] = lab;
    s_aux[lab + Kmax] = 1;                                  // value v lives at slot v + Kmax (values >= -Kmax)
  }
  __syncthreads();
  // exclusive prefix sum of the presence flags = rank of every distinct value (one warp: M + Kmax <= a few thousand)
  if (warp == 0) {
    int carry = 0;
    for (int i0 = 0; i0 < M + Kmax; i0 += 32) {
      const int i = i0 + lane;
      const int f = i < M + Kmax ? s_aux[i] : 0;
      int inc = f;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      if (i < M + Kmax) s_aux[i] = carry + inc - f;
      carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) *n_labels = carry;
  }
  __syncthreads();
  for (int i = tid; i < M; i += kSolveThreads) {
    const int dense = s_aux[s_seg[i] + Kmax];
    s_seg[i] = dense;
    labels[i] = dense;
  }
  __syncthreads();
  for (int i = tid; i < M + Kmax; i += kSolveThreads) s_aux[i] = 0;
  __syncthreads();
  for (int i = tid; i < M; i += kSolveThreads) atomicAdd(&s_aux[s_seg[i]], 1);
  __syncthreads();
  // weights[l] = 1 / (count_l + 1e-10) in float32 (one_hot / (sum + 1e-10), :48); unused tail = 0
  for (int i = tid; i < M; i += kSolveThreads)
    weights[i] = s_aux[i] > 0 ? __fdiv_rn(1.0f, __fadd_rn(static_cast<float>(s_aux[i]), 1e-10f)) : 0.f;
}

int pair_bits_for(int M) {
  int bits = 1;
  while ((1ll << bits) < static_cast<long long>(M) * M) ++bits;
  return bits;
}

struct SolveWs {
  SolveHeader *hdr;
  unsigned long long *keys, *sorted;
  void *cub_tmp;
  size_t cub_bytes, bytes;
};

SolveWs carve(void *ws, int M) {
  SolveWs w{};
  const size_t n = static_cast<size_t>(M) * (M - 1) / 2 + 1;
  size_t cub = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, cub, static_cast<const unsigned long long *>(nullptr),
                                 static_cast<unsigned long long *>(nullptr), static_cast<int>(n), 0, 64);
  unsigned char *p = static_cast<unsigned char *>(ws);
  size_t o = 0;
  w.hdr = reinterpret_cast<SolveHeader *>(p + o); o += 256;
  w.keys = reinterpret_cast<unsigned long long *>(p + o); o += (n * 8 + 255) / 256 * 256;
  w.sorted = reinterpret_cast<unsigned long long *>(p + o); o += (n * 8 + 255) / 256 * 256;
  w.cub_tmp = p + o; w.cub_bytes = cub; o += (cub + 255) / 256 * 256;
  w.bytes = o;
  return w;
}

}  // namespace
}  // namespace cpfn

using namespace cpfn;

extern "C" size_t cpfn_merge_solve_workspace_bytes(int nb, int Kl, int Kg) {
  if (nb <= 0 || Kl <= 0 || Kg <= 0) return 0;
  const long long M = static_cast<long long>(nb) * Kl + Kg;
  if (M > kMaxNodes) return 0;
  return carve(nullptr, static_cast<int>(M)).bytes;
}

extern "C" int cpfn_merge_solve(const float *similarity, int nb, int Kl, int Kg, float threshold, int32_t *labels,
                                float *label_weight, int32_t *n_labels, int32_t *segments, void *workspace,
                                size_t workspace_bytes, cpfn_stream_t stream) {
  if (!similarity || !labels || !label_weight || !n_labels || nb <= 0 || Kl <= 0 || Kg <= 0 || nb + 1 > 64)
    return CPFN_EINVAL;
  const long long Mll = static_cast<long long>(nb) * Kl + Kg;
  if (Mll > kMaxNodes) return CPFN_EINVAL;
  const int M = static_cast<int>(Mll);
  const SolveWs w = carve(workspace, M);
  if (!workspace || workspace_bytes < w.bytes || (reinterpret_cast<uintptr_t>(workspace) & 255)) return CPFN_EWORKSPACE;
  cudaStream_t s = as_stream(stream);
  const int pair_bits = pair_bits_for(M);
  solve_init_kernel<<<1, 1, 0, s>>>(w.hdr);
  const int n_slots = static_cast<int>(static_cast<long long>(M) * (M - 1) / 2 + 1);
  // the pair count stays on the device: unused key slots are all-ones and sort behind every real pair
  CPFN_CUDA_TRY(cudaMemsetAsync(w.keys, 0xff, static_cast<size_t>(n_slots) * 8, s));
  const long long entries = static_cast<long long>(M) * M;
  solve_pairs_kernel<<<static_cast<unsigned int>((entries + 255) / 256), 256, 0, s>>>(similarity, M, nb, Kl, threshold,
                                                                                     pair_bits, w.keys, w.hdr);
  size_t cub_bytes = w.cub_bytes;
  CPFN_CUDA_TRY(cub::DeviceRadixSort::SortKeys(w.cub_tmp, cub_bytes, w.keys, w.sorted, n_slots, 0, 32 + pair_bits, s));
  const int Kmax = Kl > Kg ? Kl : Kg;
  const size_t smem = static_cast<size_t>(M) * 8 + static_cast<size_t>(M) * 4 + static_cast<size_t>(M + Kmax) * 4 +
                      2 * sizeof(unsigned int) * kLoaders * kSegPairs;
  auto kern = M <= 1024 ? solve_greedy_kernel<true> : solve_greedy_kernel<false>;
  if (smem > 48 * 1024)
    CPFN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  kern<<<1, kSolveThreads, smem, s>>>(similarity, w.sorted, w.hdr, M, nb, Kl, Kg, threshold, pair_bits, labels,
                                      label_weight, n_labels, segments);
  return check_launch();
}
