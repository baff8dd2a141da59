// Patch extraction: for every seed point the k nearest points of a high-resolution cloud, ordered by
// distance -- what the reference computes per seed with numpy
//     distances = np.linalg.norm(seed - gt_points_hr, axis=1); np.argsort(distances)[:k]; np.sort(distances)[:k]
// (Utils/sampling_utils.py:9-13, Preprocessing/preprocessing_sampling_patch.py:36-40).
//
// Exact selection, no sort of the N distances: a most-significant-digit radix SELECT over the 64-bit keys
//     key = (bits(distance) << 32) | index
// (distances are >= 0, so their bit patterns order like the floats; the index in the low word makes the
// keys unique and the order the STABLE argsort order -- numpy's default introsort leaves the order of
// equal distances unspecified).  Digits 11+10+10 bits of the distance, then the index bits; a pass
// histograms its digit over the elements that match the prefix found so far, and the last CTA to finish
// (ticket) scans the histogram and extends the prefix.  As soon as the threshold bin holds exactly the
// number of elements still needed the selection is decided and the remaining passes return at once.
// The k selected keys are compacted and ordered in shared memory by one CTA per seed (bucket + rank).
//
// Distance arithmetic is numpy's for float32 rows of 3: d = sqrt((dx*dx + dy*dy) + dz*dz), every step
// rounded to fp32 (no FMA) -- bit-identical to np.linalg.norm(..., axis=1).
#include "common.cuh"

namespace cpfn {
namespace {

constexpr int kSelThreads = 512;
constexpr int kSelBinsLog = 11;
constexpr int kSelBins = 1 << kSelBinsLog;
constexpr int kSortThreads = 1024;
constexpr int kMaxPatch = 16384;

struct SelState {                 // one per seed, in the workspace
  unsigned long long prefix;      // digits decided so far (undecided low bits are 0)
  unsigned int k_rem;             // how many elements to take among those that match the prefix
  unsigned int shift;             // number of undecided low bits
  unsigned int resolved;          // 1: (key >> shift) <= (prefix >> shift) selects exactly k keys
  unsigned int ticket;            // CTAs that finished the current pass
  unsigned int n_out;             // compaction cursor
  unsigned int pad;
};

__device__ __forceinline__ unsigned int np_norm_bits(float sx, float sy, float sz, float x, float y, float z) {
  const float dx = __fsub_rn(sx, x), dy = __fsub_rn(sy, y), dz = __fsub_rn(sz, z);
  const float s = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
  return __float_as_uint(__fsqrt_rn(s));
}

__global__ void sel_init_kernel(SelState *st, unsigned int *hist, int S, int k) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < S * kSelBins) hist[i] = 0;
  if (i < S) {
    SelState s;
    s.prefix = 0; s.k_rem = static_cast<unsigned int>(k); s.shift = 64; s.resolved = 0; s.ticket = 0; s.n_out = 0; s.pad = 0;
    st[i] = s;
  }
}

__device__ __forceinline__ unsigned long long make_key(unsigned int db, int i) {
  return (static_cast<unsigned long long>(db) << 32) | static_cast<unsigned int>(i);
}

// Adds the CTA's shared histogram to a seed's global one (fire-and-forget atomics on the non-empty bins).
__device__ __forceinline__ void flush_hist(const unsigned int *s_hist, unsigned int *gh, int nb) {
  for (int i = threadIdx.x; i < nb; i += kSelThreads) {
    const unsigned int c = s_hist[i];
    if (c) atomicAdd(&gh[i], c);
  }
}

// One WARP closes a pass for one seed once every CTA has flushed: finds the bin where the running count
// reaches k_rem, extends the prefix, and leaves the global histogram zeroed and the ticket reset.
__device__ __forceinline__ void warp_close_pass(unsigned int *gh, SelState *my, unsigned long long prefix, int pos,
                                                int bits) {
  const int nb = 1 << bits, lane = lane_id();
  const int per = (nb + 31) >> 5;                               // contiguous bins per lane: 64, 32, ... or 1
  const unsigned int k_rem = my->k_rem;
  const int b0 = lane * per;
  unsigned int local = 0;
  if ((per & 15) == 0) {                                        // 16 bins per round, four independent 16-byte loads
    for (int j = 0; j < per; j += 16) {
      const uint4 *p = reinterpret_cast<const uint4 *>(gh + b0 + j);
      const uint4 a = __ldcg(p), b = __ldcg(p + 1), c = __ldcg(p + 2), e = __ldcg(p + 3);
      local += (a.x + a.y + a.z + a.w) + (b.x + b.y + b.z + b.w) + (c.x + c.y + c.z + c.w) + (e.x + e.y + e.z + e.w);
    }
  } else {
    for (int j = 0; j < per; ++j)
      if (b0 + j < nb) local += __ldcg(&gh[b0 + j]);
  }
  unsigned int incl = local;
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  const unsigned int excl = incl - local;
  if (excl < k_rem && k_rem <= incl) {                          // exactly one lane
    unsigned int run = excl;
    int found = -1;
    unsigned int found_c = 0;
    for (int j = 0; j < per && found < 0; j += 16) {
      unsigned int v[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) v[q] = (j + q < per && b0 + j + q < nb) ? __ldcg(&gh[b0 + j + q]) : 0u;
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        if (found < 0 && j + q < per && b0 + j + q < nb) {
          if (run + v[q] >= k_rem) { found = b0 + j + q; found_c = v[q]; }
          else run += v[q];
        }
      }
    }
    my->prefix = prefix | (static_cast<unsigned long long>(found) << pos);
    my->k_rem = k_rem - run;
    my->shift = static_cast<unsigned int>(pos);
    my->resolved = (found_c == k_rem - run || pos == 0) ? 1u : 0u;
    my->ticket = 0;
  }
  __syncwarp();
  for (int i = lane; i < nb; i += 32) gh[i] = 0;                // clean for the next pass / call
}

constexpr int kFirstCopies = 4;
constexpr int kFirstPts = 8;       // points per thread of the first pass: a CTA owns a tile of 4096 points

// First pass: distances + histogram of the top digit (pos 52, 11 bits: exponent and 3 mantissa bits).
// A CTA loads its tile of points ONCE into registers and loops over its share of the seeds (seeds
// blockIdx.y, blockIdx.y + gridDim.y, ...: the host splits the seeds over just enough CTA rows to fill the
// machine), so a large cloud is read once or twice whatever the number of seeds; the distance bit patterns go to dist[seed, :] (rows padded to a multiple of
// 4 with 0xffffffff, which never matches a prefix because bit 63 of a key is 0).  Most distances share a
// few bins, so a thread first merges equal bins of its own points.
__global__ void __launch_bounds__(kSelThreads, 2)
sel_first_kernel(const float *__restrict__ hr, const float *__restrict__ seeds, unsigned int *__restrict__ dist,
                 unsigned int *__restrict__ hist, SelState *__restrict__ st, int N, int Npad, int S, int aligned) {
  __shared__ unsigned int s_hist[kFirstCopies * kSelBins];    // one copy per lane & 3: fewer same-address atomics
  __shared__ unsigned int s_last[kSelThreads];
  constexpr int kChunks = kFirstPts / 4;
  unsigned int *my_hist = s_hist + (threadIdx.x & (kFirstCopies - 1)) * kSelBins;
  const int n4 = Npad >> 2;
  float v[kChunks][12];
  int chunk[kChunks];
#pragma unroll
  for (int u = 0; u < kChunks; ++u) {
    const int c = (blockIdx.x * kChunks + u) * kSelThreads + threadIdx.x;
    chunk[u] = c < n4 ? c : -1;
    const int i0 = c << 2;
    if (c < n4 && aligned && i0 + 4 <= N) {
      const float4 *p = reinterpret_cast<const float4 *>(hr + static_cast<size_t>(i0) * 3);
      const float4 a = __ldg(p), b = __ldg(p + 1), e = __ldg(p + 2);
      v[u][0] = a.x; v[u][1] = a.y; v[u][2] = a.z; v[u][3] = a.w; v[u][4] = b.x; v[u][5] = b.y; v[u][6] = b.z;
      v[u][7] = b.w; v[u][8] = e.x; v[u][9] = e.y; v[u][10] = e.z; v[u][11] = e.w;
    } else {
#pragma unroll
      for (int j = 0; j < 12; ++j)
        v[u][j] = (c < n4 && i0 + j / 3 < N) ? __ldg(hr + static_cast<size_t>(i0) * 3 + j) : 0.f;
    }
  }
  for (int seed = blockIdx.y; seed < S; seed += gridDim.y) {
    for (int i = threadIdx.x; i < kFirstCopies * kSelBins; i += kSelThreads) s_hist[i] = 0;
    __syncthreads();
    const float sx = __ldg(seeds + seed * 3), sy = __ldg(seeds + seed * 3 + 1), sz = __ldg(seeds + seed * 3 + 2);
    unsigned int *d = dist + static_cast<size_t>(seed) * Npad;
    unsigned int run_bin = 0xffffffffu, run_cnt = 0;
#pragma unroll
    for (int u = 0; u < kChunks; ++u) {
      if (chunk[u] < 0) continue;
      const int i0 = chunk[u] << 2;
      unsigned int db[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        db[j] = (i0 + j < N) ? np_norm_bits(sx, sy, sz, v[u][3 * j], v[u][3 * j + 1], v[u][3 * j + 2]) : 0xffffffffu;
      *reinterpret_cast<uint4 *>(d + i0) = make_uint4(db[0], db[1], db[2], db[3]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (i0 + j >= N) continue;
        const unsigned int bin = db[j] >> 20;                  // key bits 62..52
        if (bin == run_bin) { ++run_cnt; continue; }
        if (run_cnt) atomicAdd(&my_hist[run_bin], run_cnt);
        run_bin = bin; run_cnt = 1;
      }
    }
    if (run_cnt) atomicAdd(&my_hist[run_bin], run_cnt);
    __syncthreads();
    unsigned int *gh = hist + static_cast<size_t>(seed) * kSelBins;
    for (int i = threadIdx.x; i < kSelBins; i += kSelThreads) {
      unsigned int c = 0;
#pragma unroll
      for (int r = 0; r < kFirstCopies; ++r) c += s_hist[r * kSelBins + i];
      if (c) atomicAdd(&gh[i], c);
    }
    __syncthreads();
  }
  // tickets of all seeds at once; whichever CTA arrives last for a seed closes its pass, one warp per seed
  __threadfence();
  __syncthreads();
  const int mine = (S - static_cast<int>(blockIdx.y) + static_cast<int>(gridDim.y) - 1) / static_cast<int>(gridDim.y);
  for (int base = 0; base < mine; base += kSelThreads) {       // this row's seeds, kSelThreads at a time
    const int seed = blockIdx.y + (base + threadIdx.x) * gridDim.y;
    s_last[threadIdx.x] = (base + threadIdx.x < mine && atomicAdd(&st[seed].ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    for (int q = threadIdx.x >> 5; q < kSelThreads && base + q < mine; q += kSelThreads / 32)
      if (s_last[q]) {
        const int sq = blockIdx.y + (base + q) * gridDim.y;
        __threadfence();
        warp_close_pass(hist + static_cast<size_t>(sq) * kSelBins, st + sq, 0ull, 52, 11);
      }
    __syncthreads();
  }
}

// A later pass over digit [pos, pos+bits) of the keys of seed blockIdx.y: histogram of the elements that
// match the prefix found so far.  Eight elements per thread and iteration, both loads issued before use
// (the pass is latency-bound otherwise).
__global__ void __launch_bounds__(kSelThreads)
sel_pass_kernel(const unsigned int *__restrict__ dist, unsigned int *__restrict__ hist, SelState *__restrict__ st,
                int Npad, int pos, int bits) {
  __shared__ unsigned int s_hist[kSelBins];
  __shared__ unsigned int s_last;
  const int seed = blockIdx.y;
  SelState *my = st + seed;
  if (my->resolved) return;                                   // uniform over the whole grid row
  const unsigned int *d = dist + static_cast<size_t>(seed) * Npad;
  const unsigned long long prefix = my->prefix;
  const int nb = 1 << bits;
  for (int i = threadIdx.x; i < nb; i += kSelThreads) s_hist[i] = 0;
  __syncthreads();
  const int hi = pos + bits;                                   // bits >= hi are decided
  const unsigned int mask = static_cast<unsigned int>(nb - 1);
  const int n4 = Npad >> 2;
  const int step = gridDim.x * kSelThreads;
  const unsigned long long want = prefix >> hi;
  for (int c = blockIdx.x * kSelThreads + threadIdx.x; c < n4; c += 2 * step) {
    const uint4 a = *reinterpret_cast<const uint4 *>(d + (static_cast<size_t>(c) << 2));
    const int c2 = c + step;
    uint4 b = make_uint4(~0u, ~0u, ~0u, ~0u);
    if (c2 < n4) b = *reinterpret_cast<const uint4 *>(d + (static_cast<size_t>(c2) << 2));
    const unsigned int db[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const unsigned long long key = make_key(db[j], ((j < 4 ? c : c2) << 2) + (j & 3));
      if ((key >> hi) == want) atomicAdd(&s_hist[static_cast<unsigned int>(key >> pos) & mask], 1u);
    }
  }
  __syncthreads();
  unsigned int *gh = hist + static_cast<size_t>(seed) * kSelBins;
  flush_hist(s_hist, gh, nb);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&my->ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (s_last && threadIdx.x < 32) {
    __threadfence();
    warp_close_pass(gh, my, prefix, pos, bits);
  }
}

constexpr int kStage = 2048;       // keys a CTA stages in shared memory before one global reservation

// Selected keys -> cand[seed, 0..k) in arbitrary order.  A CTA collects its keys in shared memory and
// reserves room in the output with ONE global atomic (thousands of returning atomics on a seed's cursor
// serialise in L2); keys beyond the staging capacity fall back to one atomic each.
__global__ void __launch_bounds__(kSelThreads)
sel_compact_kernel(const unsigned int *__restrict__ dist, SelState *__restrict__ st, int Npad, int k,
                   unsigned long long *__restrict__ cand) {
  __shared__ unsigned long long s_stage[kStage];
  __shared__ unsigned int s_n, s_base;
  const int seed = blockIdx.y;
  SelState *my = st + seed;
  const unsigned int shift = my->shift;
  const unsigned long long lim = my->prefix >> shift;
  const unsigned int *d = dist + static_cast<size_t>(seed) * Npad;
  unsigned long long *out = cand + static_cast<size_t>(seed) * k;
  const int n4 = Npad >> 2;
  const int step = gridDim.x * kSelThreads;
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  for (int c = blockIdx.x * kSelThreads + threadIdx.x; c < n4; c += 2 * step) {
    const uint4 a = *reinterpret_cast<const uint4 *>(d + (static_cast<size_t>(c) << 2));
    const int c2 = c + step;
    uint4 b = make_uint4(~0u, ~0u, ~0u, ~0u);
    if (c2 < n4) b = *reinterpret_cast<const uint4 *>(d + (static_cast<size_t>(c2) << 2));
    const unsigned int db[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const unsigned long long key = make_key(db[j], ((j < 4 ? c : c2) << 2) + (j & 3));
      if ((key >> shift) <= lim) {
        const unsigned int p = atomicAdd(&s_n, 1u);
        if (p < kStage) {
          s_stage[p] = key;
        } else {
          const unsigned int slot = atomicAdd(&my->n_out, 1u);
          if (slot < static_cast<unsigned int>(k)) out[slot] = key;
        }
      }
    }
  }
  __syncthreads();
  const unsigned int n = min(s_n, static_cast<unsigned int>(kStage));
  if (threadIdx.x == 0 && n) s_base = atomicAdd(&my->n_out, n);
  __syncthreads();
  for (unsigned int i = threadIdx.x; i < n; i += kSelThreads)
    if (s_base + i < static_cast<unsigned int>(k)) out[s_base + i] = s_stage[i];
}

constexpr int kSortBuckets = 1024;
static_assert(kSortBuckets == kSortThreads, "one bucket counter per thread");

// Monotone bucket of a distance: floor(1024 (d/r)^2), r = the largest selected distance.  The number of
// points within x of a seed on a surface grows like x^2, so the buckets come out nearly even.
__device__ __forceinline__ int sort_bucket(unsigned int dbits, float inv_r) {
  const float t = __fmul_rn(__uint_as_float(dbits), inv_r);
  const int b = static_cast<int>(__fmul_rn(__fmul_rn(t, t), static_cast<float>(kSortBuckets)));
  return b < kSortBuckets - 1 ? b : kSortBuckets - 1;
}

// One CTA per seed orders the k selected keys: counting sort into 1024 monotone distance buckets in shared
// memory, then every key is ranked inside its bucket by counting the smaller keys (keys are unique).
// Typical buckets hold ~k/1024 keys; a degenerate cloud (all distances equal) lands in one bucket and is
// still ranked exactly, only slowly.
__global__ void __launch_bounds__(kSortThreads)
sel_sort_kernel(const unsigned long long *__restrict__ cand, int k, int32_t *__restrict__ out_idx,
                float *__restrict__ out_dist, float *__restrict__ out_radius) {
  extern __shared__ unsigned long long s_key[];               // k keys, bucket by bucket
  __shared__ unsigned int s_cnt[kSortBuckets], s_start[kSortBuckets], s_cur[kSortBuckets], s_warp[kSortThreads / 32];
  __shared__ unsigned int s_max;
  const int seed = blockIdx.x, t = threadIdx.x;
  const unsigned long long *in = cand + static_cast<size_t>(seed) * k;
  if (t == 0) s_max = 0;
  s_cnt[t] = 0;                                               // kSortBuckets == kSortThreads
  __syncthreads();
  unsigned int m = 0;
  for (int i = t; i < k; i += kSortThreads) m = max(m, static_cast<unsigned int>(in[i] >> 32));
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((t & 31) == 0) atomicMax(&s_max, m);
  __syncthreads();
  const float r = __uint_as_float(s_max);
  const float inv_r = r > 0.f ? __frcp_rn(r) : 0.f;
  for (int i = t; i < k; i += kSortThreads) atomicAdd(&s_cnt[sort_bucket(static_cast<unsigned int>(in[i] >> 32), inv_r)], 1u);
  __syncthreads();
  // exclusive scan of the 1024 counts, one per thread
  const unsigned int c = s_cnt[t];
  unsigned int incl = c;
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int v = __shfl_up_sync(0xffffffffu, incl, o);
    if ((t & 31) >= o) incl += v;
  }
  if ((t & 31) == 31) s_warp[t >> 5] = incl;
  __syncthreads();
  if (t < 32) {
    unsigned int w = s_warp[t], wi = w;
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int v = __shfl_up_sync(0xffffffffu, wi, o);
      if (t >= o) wi += v;
    }
    s_warp[t] = wi - w;
  }
  __syncthreads();
  const unsigned int start = s_warp[t >> 5] + incl - c;
  s_start[t] = start;
  s_cur[t] = start;
  __syncthreads();
  for (int i = t; i < k; i += kSortThreads) {
    const unsigned long long key = in[i];
    s_key[atomicAdd(&s_cur[sort_bucket(static_cast<unsigned int>(key >> 32), inv_r)], 1u)] = key;
  }
  __syncthreads();
  for (int i = t; i < k; i += kSortThreads) {
    const unsigned long long key = s_key[i];
    const int b = sort_bucket(static_cast<unsigned int>(key >> 32), inv_r);
    const unsigned int lo = s_start[b], hi = lo + s_cnt[b];
    unsigned int rank = lo;
    for (unsigned int q = lo; q < hi; ++q) rank += s_key[q] < key ? 1u : 0u;
    out_idx[static_cast<size_t>(seed) * k + rank] = static_cast<int32_t>(key & 0xffffffffu);
    if (out_dist) out_dist[static_cast<size_t>(seed) * k + rank] = __uint_as_float(static_cast<unsigned int>(key >> 32));
  }
  if (out_radius && t == 0) out_radius[seed] = r;
}

size_t align256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

}  // namespace
}  // namespace cpfn

using namespace cpfn;

extern "C" size_t cpfn_extract_patches_workspace_bytes(int N, int S, int k) {
  if (N <= 0 || S <= 0 || k <= 0) return 0;
  return align256(sizeof(SelState) * S) + align256(sizeof(unsigned int) * kSelBins * S) +
         align256(sizeof(unsigned int) * static_cast<size_t>((N + 3) & ~3) * S) +
         align256(sizeof(unsigned long long) * static_cast<size_t>(k) * S);
}

extern "C" int cpfn_extract_patches(const float *hr_xyz, int N, const float *seeds_xyz, int S, int k, int32_t *out_idx,
                                    float *out_dist, float *out_radius, void *workspace, size_t workspace_bytes,
                                    cpfn_stream_t stream) {
  if (!hr_xyz || !seeds_xyz || !out_idx || N <= 0 || S <= 0 || S > 65535 || k <= 0 || k > N || k > kMaxPatch)
    return CPFN_EINVAL;
  if (!workspace || workspace_bytes < cpfn_extract_patches_workspace_bytes(N, S, k) ||
      (reinterpret_cast<uintptr_t>(workspace) & 15))
    return CPFN_EWORKSPACE;
  cudaStream_t s = as_stream(stream);
  char *w = static_cast<char *>(workspace);
  SelState *st = reinterpret_cast<SelState *>(w);           w += align256(sizeof(SelState) * S);
  unsigned int *hist = reinterpret_cast<unsigned int *>(w); w += align256(sizeof(unsigned int) * kSelBins * S);
  const int Npad = (N + 3) & ~3;
  unsigned int *dist = reinterpret_cast<unsigned int *>(w); w += align256(sizeof(unsigned int) * static_cast<size_t>(Npad) * S);
  unsigned long long *cand = reinterpret_cast<unsigned long long *>(w);

  const int sms = sm_count();
  if (sms <= 0) return CPFN_ELAUNCH;
  // CTAs per seed: enough to fill the machine across the S seeds, at least 8 elements per thread
  int per_seed = (4 * sms + S - 1) / S;
  const int need = (N + 8 * kSelThreads - 1) / (8 * kSelThreads);
  const int aligned = (reinterpret_cast<uintptr_t>(hr_xyz) & 15) == 0 ? 1 : 0;
  if (per_seed > need) per_seed = need;
  if (per_seed < 1) per_seed = 1;
  const dim3 grid(per_seed, S);

  sel_init_kernel<<<(S * kSelBins + 255) / 256, 256, 0, s>>>(st, hist, S, k);
  // distance bits 62..32 of the key (bit 63, the sign, is always 0): 11 + 10 + 10
  const int tiles = (Npad / 4 + kSelThreads * (kFirstPts / 4) - 1) / (kSelThreads * (kFirstPts / 4));
  int rows = (2 * sms + tiles - 1) / tiles;                    // seed rows: enough CTAs for two per SM
  if (rows > S) rows = S;
  sel_first_kernel<<<dim3(tiles, rows), kSelThreads, 0, s>>>(hr_xyz, seeds_xyz, dist, hist, st, N, Npad, S, aligned);
  sel_pass_kernel<<<grid, kSelThreads, 0, s>>>(dist, hist, st, Npad, 42, 10);
  sel_pass_kernel<<<grid, kSelThreads, 0, s>>>(dist, hist, st, Npad, 32, 10);
  // index bits (only run when equal distances straddle the cut)
  int nb = 1;
  while ((1ll << nb) < N) ++nb;
  for (int top = nb; top > 0;) {
    const int bits = top > 11 ? (top + 1) / 2 > 11 ? 11 : (top + 1) / 2 : top;
    top -= bits;
    sel_pass_kernel<<<grid, kSelThreads, 0, s>>>(dist, hist, st, Npad, top, bits);
  }
  sel_compact_kernel<<<grid, kSelThreads, 0, s>>>(dist, st, Npad, k, cand);
  const size_t smem = sizeof(unsigned long long) * k;
  static thread_local int attr_dev = -1;
  int dev = 0;
  CPFN_CUDA_TRY(cudaGetDevice(&dev));
  if (attr_dev != dev) {
    CPFN_CUDA_TRY(cudaFuncSetAttribute(sel_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(sizeof(unsigned long long) * kMaxPatch)));
    attr_dev = dev;
  }
  sel_sort_kernel<<<S, kSortThreads, smem, s>>>(cand, k, out_idx, out_dist, out_radius);
  return check_launch();
}
