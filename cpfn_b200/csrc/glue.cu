// Small fused helpers of the inference forward (each replaces a chain of tiny element-wise
// torch kernels of the reference's Python modules):
//   cpfn_three_nn_weights  three_nn + sqrt + inverse-distance weights
//                          (modules/geometry_utils.py:184, pointset_feature_propagation.py:38-42)
//   cpfn_gather_xyz        centroid gather new_xyz = xyz[fps_idx] (pointset_abstraction.py:50)
//   cpfn_spfn_post         X = normalize(head0), W = softmax(head2) (Utils/training_utils.py:141-142)
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "nn_grid.cuh"

namespace cpfn {
namespace {

constexpr int kGlueThreads = 256;
constexpr int kNnTile = 2048;

// Same scan as three_nn_kernel (interp_group.cu; bit-exact indices and squared distances), then
// d = sqrt(d2); r = 1 / (d + 1e-8); w = r / ((r0 + r1) + r2)   -- every step correctly rounded fp32,
// the order torch's element-wise kernels apply them in the reference.
__global__ void __launch_bounds__(kGlueThreads)
three_nn_weights_kernel(const float *__restrict__ unknown, const float *__restrict__ known, int n, int m,
                        float *__restrict__ weight, int32_t *__restrict__ idx) {
  extern __shared__ float4 s_known[];      // min(m, kNnTile) records: small enough to co-reside with the MLP chains
  const int b = blockIdx.y;
  const int j = blockIdx.x * kGlueThreads + threadIdx.x;
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
    for (int k = threadIdx.x; k < tn; k += kGlueThreads) {
      const float *s = kn + static_cast<size_t>(t0 + k) * 3;
      s_known[k] = make_float4(__ldg(s), __ldg(s + 1), __ldg(s + 2), 0.f);
    }
    __syncthreads();
#pragma unroll 4
    for (int k = 0; k < tn; ++k) {
      const float4 c = s_known[k];
      const float d = sqdist3(ux, uy, uz, c.x, c.y, c.z);
      if (d < b3) {
        const int kk = t0 + k;
        if (d < b1) { b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = kk; }
        else if (d < b2) { b3 = b2; i3 = i2; b2 = d; i2 = kk; }
        else { b3 = d; i3 = kk; }
      }
    }
  }
  if (j < n) {
    const float r1 = __frcp_rn(__fadd_rn(__fsqrt_rn(b1), 1e-8f));
    const float r2 = __frcp_rn(__fadd_rn(__fsqrt_rn(b2), 1e-8f));
    const float r3 = __frcp_rn(__fadd_rn(__fsqrt_rn(b3), 1e-8f));
    const float s = __fadd_rn(__fadd_rn(r1, r2), r3);
    const size_t o = (static_cast<size_t>(b) * n + j) * 3;
    weight[o] = __fdiv_rn(r1, s); weight[o + 1] = __fdiv_rn(r2, s); weight[o + 2] = __fdiv_rn(r3, s);
    idx[o] = i1; idx[o + 1] = i2; idx[o + 2] = i3;
  }
}

// The same result through the shared-memory grid of nn_grid.cuh (known clouds of <= kNnGridMax points): ~50
// candidates per query instead of all m.  Every CTA builds the grid once and answers kNnGridQ queries per thread.
constexpr int kNnGridQ = 2;

// `sorted_q` (optional): the queries as (x, y, z, index bits) records in a spatially sorted order -- the cell-ordered
// copy of the cloud that the ball-query grid of the same cloud already holds (cpfn_ball_query_grid_build).  Thread j
// then answers record j and writes the result at the record's own index: the 32 queries of a warp sit in one or two
// cells of that grid, walk the same cells of the known cloud's grid and load the same records (broadcasts) instead of
// diverging over the whole cloud.  Same result, query by query.
__global__ void __launch_bounds__(kGlueThreads)
three_nn_weights_grid_kernel(const float *__restrict__ unknown, const float4 *__restrict__ sorted_q,
                             const float *__restrict__ known, int n, int m,
                             float *__restrict__ weight, int32_t *__restrict__ idx) {
  extern __shared__ __align__(16) unsigned char s_grid_raw[];
  const int b = blockIdx.y;
  NnGrid g;
  float4 *recs;
  int *cell_start;
  nn_grid_build(known + static_cast<size_t>(b) * m * 3, m, s_grid_raw, g, recs, cell_start);
#pragma unroll 1
  for (int q = 0; q < kNnGridQ; ++q) {
    int j = (blockIdx.x * kNnGridQ + q) * kGlueThreads + threadIdx.x;
    if (j >= n) continue;
    float ux, uy, uz;
    if (sorted_q != nullptr) {
      const float4 rec = __ldg(sorted_q + static_cast<size_t>(b) * n + j);
      ux = rec.x; uy = rec.y; uz = rec.z;
      j = __float_as_int(rec.w);
    } else {
      const float *u = unknown + (static_cast<size_t>(b) * n + j) * 3;
      ux = __ldg(u); uy = __ldg(u + 1); uz = __ldg(u + 2);
    }
    const Nn3 r = nn_grid_query(ux, uy, uz, g, recs, cell_start);
    const float r1 = __frcp_rn(__fadd_rn(__fsqrt_rn(r.d1), 1e-8f));
    const float r2 = __frcp_rn(__fadd_rn(__fsqrt_rn(r.d2), 1e-8f));
    const float r3 = __frcp_rn(__fadd_rn(__fsqrt_rn(r.d3), 1e-8f));
    const float s = __fadd_rn(__fadd_rn(r1, r2), r3);
    const size_t o = (static_cast<size_t>(b) * n + j) * 3;
    weight[o] = __fdiv_rn(r1, s); weight[o + 1] = __fdiv_rn(r2, s); weight[o + 2] = __fdiv_rn(r3, s);
    idx[o] = r.i1; idx[o + 1] = r.i2; idx[o + 2] = r.i3;
  }
}

__global__ void __launch_bounds__(kGlueThreads)
gather_xyz_kernel(const float *__restrict__ xyz, const int32_t *__restrict__ idx, int N, int S, long long total,
                  float *__restrict__ out) {
  const long long e = static_cast<long long>(blockIdx.x) * kGlueThreads + threadIdx.x;   // over B*S*3
  if (e >= total) return;
  const long long row = e / 3;
  const int c = static_cast<int>(e - row * 3);
  const long long b = row / S;
  out[e] = __ldg(xyz + (b * N + __ldg(idx + row)) * 3 + c);
}

// One thread per point; the block's rows are staged through shared memory so that both the read of
// heads [rows, ld] and the writes of X [rows,3] / W [rows,K] are coalesced (ld and K are odd multiples
// of a word in practice -- 35 and 28 -- so the per-thread strided shared-memory accesses are
// conflict-free or 4-way at worst).  X = x / max(||x||, 1e-12) (F.normalize, p=2); W = softmax(logits).
template <int KMAX>
__global__ void __launch_bounds__(kGlueThreads)
spfn_post_kernel(const float *__restrict__ heads, long long rows, int ld, int x_off, int t_off, int n_types,
                 int w_off, int K, float *__restrict__ X, float *__restrict__ W, int32_t *__restrict__ inst,
                 int32_t *__restrict__ type, float *__restrict__ T_out, int rows_per_patch, int patch_stride,
                 int patch_offset) {
  extern __shared__ float s_rows[];          // [kGlueThreads * max(ld, K)]
  const long long r0 = static_cast<long long>(blockIdx.x) * kGlueThreads;
  const int nr = static_cast<int>(min(static_cast<long long>(kGlueThreads), rows - r0));
  // destination rows of X / W / T_out: identity, or -- patch-sharded cascade -- local patch j goes to the slab of
  // patch j * patch_stride + patch_offset of the (possibly peer-mapped) destination (blocks never straddle patches)
  long long q0 = r0;
  if (rows_per_patch > 0) {
    const long long j = r0 / rows_per_patch;
    q0 = (j * patch_stride + patch_offset) * rows_per_patch + (r0 - j * rows_per_patch);
  }
  const float *src = heads + r0 * ld;
  for (int i = threadIdx.x; i < nr * ld; i += kGlueThreads) s_rows[i] = __ldg(src + i);
  __syncthreads();
  if (T_out != nullptr) {                      // the type logits as their own [rows, n_types] array
    float *dt = T_out + q0 * n_types;
    for (int i = threadIdx.x; i < nr * n_types; i += kGlueThreads) {
      const int r = i / n_types;
      dt[i] = s_rows[r * ld + t_off + (i - r * n_types)];
    }
  }
  float v[KMAX];
  float xn = 0.f, yn = 0.f, zn = 0.f;
  if (threadIdx.x < nr) {
    const float *h = s_rows + threadIdx.x * ld;
    const float x = h[x_off], y = h[x_off + 1], z = h[x_off + 2];
    const float nrm = fmaxf(sqrtf(x * x + y * y + z * z), 1e-12f);
    xn = x / nrm; yn = y / nrm; zn = z / nrm;
    float mx = -INFINITY;
    int amax = 0;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      v[k] = k < K ? h[w_off + k] : -INFINITY;
      if (v[k] > mx) { mx = v[k]; amax = k; }          // first maximum, like torch.argmax
    }
    if (inst != nullptr) inst[r0 + threadIdx.x] = amax;
    if (type != nullptr) {
      float tm = -INFINITY;
      int ta = 0;
      for (int k = 0; k < n_types; ++k) {
        const float tv = h[t_off + k];
        if (tv > tm) { tm = tv; ta = k; }
      }
      type[r0 + threadIdx.x] = ta;
    }
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      v[k] = k < K ? expf(v[k] - mx) : 0.f;
      sum += v[k];
    }
    const float inv = 1.0f / sum;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) v[k] = v[k] * inv;
  }
  __syncthreads();
  if (threadIdx.x < nr) {
    float *o = s_rows + threadIdx.x * K;
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
      if (k < K) o[k] = v[k];
    X[(q0 + threadIdx.x) * 3] = xn; X[(q0 + threadIdx.x) * 3 + 1] = yn; X[(q0 + threadIdx.x) * 3 + 2] = zn;
  }
  __syncthreads();
  float *dst = W + q0 * K;
  for (int i = threadIdx.x; i < nr * K; i += kGlueThreads) dst[i] = s_rows[i];
}

// out[r, co] = bias[co] + sum_k W[co, k] * x[r, k]  (fp32, one warp per output): the per-cloud
// constant part of FP1's first layer (the broadcast global feature, 16 rows x 1024 -> 256).
__global__ void __launch_bounds__(kGlueThreads)
linear_rows_kernel(const float *__restrict__ x, const float *__restrict__ W, const float *__restrict__ bias,
                   int rows, int cin, int cout, int ldo, float *__restrict__ out) {
  const int warp = (blockIdx.x * kGlueThreads + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows * cout) return;
  const int r = warp / cout, co = warp - r * cout;
  const float *xr = x + static_cast<size_t>(r) * cin, *w = W + static_cast<size_t>(co) * cin;
  float acc = 0.f;
  for (int k = lane; k < cin; k += 32) acc = fmaf(__ldg(w + k), __ldg(xr + k), acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) out[static_cast<size_t>(r) * ldo + co] = acc + __ldg(bias + co);
  if (co == 0)                                  // zero padding of the row (columns cout .. ldo-1)
    for (int k = cout + lane; k < ldo; k += 32) out[static_cast<size_t>(r) * ldo + k] = 0.f;
}

// ---- dropout keep-mask as bits, drawn from torch's own Philox stream ----------------------------
// Philox4x32-10 as curand implements it (curand_philox4x32_x.h; torch's fused_dropout_kernel_vec calls
// curand_init(seed, thread, offset) + curand_uniform4 per 4 consecutive elements).
__device__ __forceinline__ uint4 philox_round(uint4 c, uint2 k) {
  const unsigned int hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
  const unsigned int hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
  return make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
}
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 9; ++r) {
    c = philox_round(c, k);
    k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
  }
  return philox_round(c, k);
}
// curand_uniform: x * 2^-32 + 2^-33 in fp32, in (0, 1]
__device__ __forceinline__ bool keep_bit(unsigned int x, float keep) {
  return __fmaf_rn(__uint2float_rn(x), 2.3283064e-10f, 1.16415322e-10f) < keep;
}

// Per-patch normalisation of the LocalSPFN input (Dataset/dataloaders.py:249-253): gather the patch's points, centre
// them on their mean, scale by the largest distance from it.  One CTA per patch; the mean is accumulated in fp64 in
// a FIXED order (thread-strided partial sums, shuffle tree, warp 0 over the warp sums), so a patch's result does not
// depend on how many patches share the call -- torch's reductions pick their strategy from the tensor shape.
template <typename IndexT>
__global__ void __launch_bounds__(1024)
normalise_patches_kernel(const float *__restrict__ P, const IndexT *__restrict__ idx, int Np, long long Ng,
                         float *__restrict__ out) {
  __shared__ double s_sum[32][3];
  __shared__ float s_mean[3], s_max[32], s_scale;
  const IndexT *my = idx + static_cast<size_t>(blockIdx.x) * Np;
  float *o = out + static_cast<size_t>(blockIdx.x) * Np * 3;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double sx = 0.0, sy = 0.0, sz = 0.0;
  for (int i = threadIdx.x; i < Np; i += 1024) {
    const long long g = static_cast<long long>(my[i]);
    const float *p = P + (g >= 0 && g < Ng ? g : 0) * 3;
    sx += static_cast<double>(__ldg(p)); sy += static_cast<double>(__ldg(p + 1)); sz += static_cast<double>(__ldg(p + 2));
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    sx += __shfl_down_sync(0xffffffffu, sx, off);
    sy += __shfl_down_sync(0xffffffffu, sy, off);
    sz += __shfl_down_sync(0xffffffffu, sz, off);
  }
  if (lane == 0) { s_sum[warp][0] = sx; s_sum[warp][1] = sy; s_sum[warp][2] = sz; }
  __syncthreads();
  if (warp == 0) {
    double a = s_sum[lane][0], b = s_sum[lane][1], c = s_sum[lane][2];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      a += __shfl_down_sync(0xffffffffu, a, off);
      b += __shfl_down_sync(0xffffffffu, b, off);
      c += __shfl_down_sync(0xffffffffu, c, off);
    }
    if (lane == 0) {
      s_mean[0] = static_cast<float>(a / Np); s_mean[1] = static_cast<float>(b / Np); s_mean[2] = static_cast<float>(c / Np);
    }
  }
  __syncthreads();
  const float mx = s_mean[0], my_ = s_mean[1], mz = s_mean[2];
  float best = 0.f;
  for (int i = threadIdx.x; i < Np; i += 1024) {
    const long long g = static_cast<long long>(my[i]);
    const float *p = P + (g >= 0 && g < Ng ? g : 0) * 3;
    const float dx = __fsub_rn(__ldg(p), mx), dy = __fsub_rn(__ldg(p + 1), my_), dz = __fsub_rn(__ldg(p + 2), mz);
    best = fmaxf(best, __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz))));
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, off));
  if (lane == 0) s_max[warp] = best;
  __syncthreads();
  if (warp == 0) {
    float m = s_max[lane];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    if (lane == 0) s_scale = m;
  }
  __syncthreads();
  const float scale = s_scale;
  for (int i = threadIdx.x; i < Np; i += 1024) {
    const long long g = static_cast<long long>(my[i]);
    const float *p = P + (g >= 0 && g < Ng ? g : 0) * 3;
    o[i * 3] = __fdiv_rn(__fsub_rn(__ldg(p), mx), scale);
    o[i * 3 + 1] = __fdiv_rn(__fsub_rn(__ldg(p + 1), my_), scale);
    o[i * 3 + 2] = __fdiv_rn(__fsub_rn(__ldg(p + 2), mz), scale);
  }
}

__global__ void rng_set_kernel(unsigned long long *state, unsigned long long seed, unsigned long long offset) {
  state[0] = seed;
  state[1] = offset;
}

// One thread per (4 consecutive rows, 32-channel word).  Element (b, c, n) of the channel-major tensor has linear
// index e = (b*C + c)*N + n; torch's thread (e/4) % T draws it in iteration (e/4) / T as component e % 4.
// FAST: N % 4 == 0 and fewer than 2^32 elements -- the four rows of a thread share one Philox block per channel
// and all index arithmetic is 32-bit (the division by T through a float estimate with an exact fix-up).
template <bool FAST>
__global__ void __launch_bounds__(kGlueThreads)
dropout_bits_kernel(const unsigned long long *__restrict__ state, int C, int N, long long rows, float keep,
                    unsigned int T, int words, uint32_t *__restrict__ bits) {
  const long long t = static_cast<long long>(blockIdx.x) * kGlueThreads + threadIdx.x;
  const long long quad = t / words;
  const int cw = static_cast<int>(t - quad * words);
  const long long r0 = quad * 4;
  if (r0 >= rows) return;
  const unsigned long long seed = state[0], off4 = state[1] >> 2;
  // the key schedule is the same for every block: ten (x, y) pairs, computed once
  uint2 ks[10];
  ks[0] = make_uint2(static_cast<unsigned int>(seed), static_cast<unsigned int>(seed >> 32));
#pragma unroll
  for (int r = 1; r < 10; ++r) ks[r] = make_uint2(ks[r - 1].x + 0x9E3779B9u, ks[r - 1].y + 0xBB67AE85u);
  auto philox = [&](unsigned long long ctr, unsigned int sub) {
    uint4 c = make_uint4(static_cast<unsigned int>(ctr), static_cast<unsigned int>(ctr >> 32), sub, 0u);
#pragma unroll
    for (int r = 0; r < 10; ++r) c = philox_round(c, ks[r]);
    return c;
  };
  const int c_end = min(C, cw * 32 + 32);
  uint32_t w[4] = {0u, 0u, 0u, 0u};
  if (FAST) {
    const unsigned int b = static_cast<unsigned int>(r0 / N);
    const unsigned int q0 = (b * static_cast<unsigned int>(C) * static_cast<unsigned int>(N) +
                             static_cast<unsigned int>(r0 - static_cast<long long>(b) * N)) >> 2;   // e/4 at channel 0
    const unsigned int qstep = static_cast<unsigned int>(N) >> 2;
    const float inv_t = 1.0f / static_cast<float>(T);
    for (int c = cw * 32; c < c_end; ++c) {
      const unsigned int q = q0 + static_cast<unsigned int>(c) * qstep;
      unsigned int k = static_cast<unsigned int>(static_cast<float>(q) * inv_t);
      int sub = static_cast<int>(q - k * T);
      if (sub < 0) { sub += static_cast<int>(T); --k; }
      else if (sub >= static_cast<int>(T)) { sub -= static_cast<int>(T); ++k; }
      const uint4 o = philox(off4 + k, static_cast<unsigned int>(sub));
      const unsigned int bit = 1u << (c & 31);
      if (keep_bit(o.x, keep)) w[0] |= bit;
      if (keep_bit(o.y, keep)) w[1] |= bit;
      if (keep_bit(o.z, keep)) w[2] |= bit;
      if (keep_bit(o.w, keep)) w[3] |= bit;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) bits[(r0 + i) * words + cw] = w[i];
    return;
  }
  unsigned long long base[4];
  int nr = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long r = r0 + i;
    if (r < rows) {
      const long long b = r / N;
      base[i] = static_cast<unsigned long long>(b) * C * N + static_cast<unsigned long long>(r - b * N);
      nr = i + 1;
    } else {
      base[i] = 0;
    }
  }
  for (int c = cw * 32; c < c_end; ++c) {
    unsigned long long have = ~0ull;
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (i < nr) {
        const unsigned long long e = base[i] + static_cast<unsigned long long>(c) * N;
        const unsigned long long q = e >> 2;
        if (q != have) {
          unsigned long long k;
          if ((q >> 32) == 0) k = static_cast<unsigned int>(q) / T;       // 32-bit division in the common case
          else k = q / T;
          o = philox(off4 + k, static_cast<unsigned int>(q - k * T));
          have = q;
        }
        const int comp = static_cast<int>(e & 3);
        const unsigned int x = comp == 0 ? o.x : (comp == 1 ? o.y : (comp == 2 ? o.z : o.w));
        if (keep_bit(x, keep)) w[i] |= 1u << (c & 31);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (i < nr) bits[(r0 + i) * words + cw] = w[i];
}

}  // namespace
}  // namespace cpfn

extern "C" int cpfn_linear_rows(const float *x, const float *W, const float *bias, int rows, int cin, int cout,
                                int ldo, float *out, cpfn_stream_t stream) {
  using namespace cpfn;
  if (rows < 0 || cin <= 0 || cout <= 0 || ldo < cout) return CPFN_EINVAL;
  if (rows == 0) return CPFN_OK;
  if (!x || !W || !bias || !out) return CPFN_EINVAL;
  const long long warps = static_cast<long long>(rows) * cout;
  const unsigned grid = static_cast<unsigned>((warps * 32 + kGlueThreads - 1) / kGlueThreads);
  linear_rows_kernel<<<grid, kGlueThreads, 0, as_stream(stream)>>>(x, W, bias, rows, cin, cout, ldo, out);
  return check_launch();
}

extern "C" int cpfn_normalise_patches(const float *points, long long Ng, const void *patch_idx, int idx_is_int64,
                                      int nb, int Np, float *out, cpfn_stream_t stream) {
  using namespace cpfn;
  if (nb < 0 || Np <= 0 || Ng <= 0) return CPFN_EINVAL;
  if (nb == 0) return CPFN_OK;
  if (!points || !patch_idx || !out) return CPFN_EINVAL;
  if (idx_is_int64)
    normalise_patches_kernel<long long><<<nb, 1024, 0, as_stream(stream)>>>(points, static_cast<const long long *>(patch_idx), Np, Ng, out);
  else
    normalise_patches_kernel<int32_t><<<nb, 1024, 0, as_stream(stream)>>>(points, static_cast<const int32_t *>(patch_idx), Np, Ng, out);
  return check_launch();
}

extern "C" int cpfn_zero_fill(void *dst, size_t bytes, cpfn_stream_t stream) {
  using namespace cpfn;
  if (bytes == 0) return CPFN_OK;
  if (!dst) return CPFN_EINVAL;
  CPFN_CUDA_TRY(cudaMemsetAsync(dst, 0, bytes, as_stream(stream)));
  return CPFN_OK;
}

extern "C" int cpfn_rng_set(unsigned long long *rng_state, unsigned long long seed, unsigned long long offset,
                            cpfn_stream_t stream) {
  using namespace cpfn;
  if (!rng_state) return CPFN_EINVAL;
  rng_set_kernel<<<1, 1, 0, as_stream(stream)>>>(rng_state, seed, offset);
  return check_launch();
}

extern "C" int cpfn_dropout_mask_bits(const unsigned long long *rng_state, int B, int C, int N, float keep_prob,
                                      long long torch_threads, uint32_t *bits, cpfn_stream_t stream) {
  using namespace cpfn;
  if (B < 0 || C <= 0 || N < 0 || torch_threads <= 0 || torch_threads > 0xFFFFFFFFll) return CPFN_EINVAL;
  if (B == 0 || N == 0) return CPFN_OK;
  if (!rng_state || !bits) return CPFN_EINVAL;
  const int words = (C + 31) / 32;
  const long long rows = static_cast<long long>(B) * N;
  const long long threads = ((rows + 3) / 4) * words;
  const long long grid = (threads + kGlueThreads - 1) / kGlueThreads;
  if (grid > 0x7FFFFFFFll) return CPFN_EINVAL;
  const bool fast = (N % 4) == 0 && rows * C < 0xFFFFFFFFll && torch_threads < 0x7FFFFFFFll &&
                    rows * C / 4 < (1ll << 24) * 64;   // float estimate of q / T stays within one of the quotient
  if (fast)
    dropout_bits_kernel<true><<<static_cast<unsigned>(grid), kGlueThreads, 0, as_stream(stream)>>>(
        rng_state, C, N, rows, keep_prob, static_cast<unsigned int>(torch_threads), words, bits);
  else
    dropout_bits_kernel<false><<<static_cast<unsigned>(grid), kGlueThreads, 0, as_stream(stream)>>>(
        rng_state, C, N, rows, keep_prob, static_cast<unsigned int>(torch_threads), words, bits);
  return check_launch();
}

extern "C" int cpfn_three_nn_weights(const float *unknown, const float *known, int B, int n, int m,
                                     float *weight, int32_t *idx, cpfn_stream_t stream) {
  using namespace cpfn;
  if (B < 0 || n < 0 || m < 0) return CPFN_EINVAL;
  if (B == 0 || n == 0) return CPFN_OK;
  if (!unknown || !weight || !idx || (m > 0 && !known) || B > 65535) return CPFN_EINVAL;
  if (m >= kNnGridMin && m <= kNnGridMax && getenv("CPFN_NN_NO_GRID") == nullptr) {
    const size_t gsmem = nn_grid_smem_bytes(m);
    if (gsmem > 48 * 1024)
      CPFN_CUDA_TRY(cudaFuncSetAttribute(three_nn_weights_grid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(gsmem)));
    dim3 ggrid((n + kGlueThreads * kNnGridQ - 1) / (kGlueThreads * kNnGridQ), B);
    three_nn_weights_grid_kernel<<<ggrid, kGlueThreads, gsmem, as_stream(stream)>>>(unknown, nullptr, known, n, m, weight, idx);
    return check_launch();
  }
  dim3 grid((n + kGlueThreads - 1) / kGlueThreads, B);
  const size_t smem = sizeof(float4) * static_cast<size_t>(m < kNnTile ? (m > 0 ? m : 1) : kNnTile);
  three_nn_weights_kernel<<<grid, kGlueThreads, smem, as_stream(stream)>>>(unknown, known, n, m, weight, idx);
  return check_launch();
}

extern "C" int cpfn_three_nn_weights_sorted(const void *sorted_queries, const float *known, int B, int n, int m,
                                            float *weight, int32_t *idx, cpfn_stream_t stream) {
  using namespace cpfn;
  if (B < 0 || n < 0 || m < 0) return CPFN_EINVAL;
  if (B == 0 || n == 0) return CPFN_OK;
  if (!sorted_queries || !known || !weight || !idx || B > 65535 || m < kNnGridMin || m > kNnGridMax) return CPFN_EINVAL;
  const size_t gsmem = nn_grid_smem_bytes(m);
  if (gsmem > 48 * 1024)
    CPFN_CUDA_TRY(cudaFuncSetAttribute(three_nn_weights_grid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(gsmem)));
  dim3 ggrid((n + kGlueThreads * kNnGridQ - 1) / (kGlueThreads * kNnGridQ), B);
  three_nn_weights_grid_kernel<<<ggrid, kGlueThreads, gsmem, as_stream(stream)>>>(
      nullptr, static_cast<const float4 *>(sorted_queries), known, n, m, weight, idx);
  return check_launch();
}

extern "C" int cpfn_gather_xyz(const float *xyz, const int32_t *idx, int B, int N, int S, float *out,
                               cpfn_stream_t stream) {
  using namespace cpfn;
  if (B < 0 || N < 0 || S < 0) return CPFN_EINVAL;
  if (B == 0 || S == 0) return CPFN_OK;
  if (!xyz || !idx || !out || N == 0) return CPFN_EINVAL;
  const long long total = static_cast<long long>(B) * S * 3;
  gather_xyz_kernel<<<static_cast<unsigned>((total + kGlueThreads - 1) / kGlueThreads), kGlueThreads, 0,
                      as_stream(stream)>>>(xyz, idx, N, S, total, out);
  return check_launch();
}

extern "C" int cpfn_spfn_post_scatter(const float *heads, long long rows, int ld, int x_off, int t_off, int n_types,
                                      int w_off, int K, float *X, float *W, float *T_out, int32_t *inst, int32_t *type,
                                      int rows_per_patch, int patch_stride, int patch_offset, cpfn_stream_t stream) {
  using namespace cpfn;
  if (rows < 0 || K <= 0 || K > 64 || ld < 3 || x_off < 0 || w_off < 0 || x_off + 3 > ld || w_off + K > ld ||
      ((type != nullptr || T_out != nullptr) && (t_off < 0 || n_types <= 0 || t_off + n_types > ld))) return CPFN_EINVAL;
  if (rows_per_patch < 0 || (rows_per_patch > 0 && ((rows_per_patch % kGlueThreads) != 0 || patch_stride <= 0 ||
                                                     patch_offset < 0 || patch_offset >= patch_stride)))
    return CPFN_EINVAL;
  if (rows == 0) return CPFN_OK;
  if (!heads || !X || !W) return CPFN_EINVAL;
  const unsigned grid = static_cast<unsigned>((rows + kGlueThreads - 1) / kGlueThreads);
  if (ld > 128) return CPFN_EINVAL;
  const size_t smem = static_cast<size_t>(kGlueThreads) * (ld > K ? ld : K) * sizeof(float);
  if (K <= 32) {
    if (smem > 48 * 1024) CPFN_CUDA_TRY(cudaFuncSetAttribute(spfn_post_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    spfn_post_kernel<32><<<grid, kGlueThreads, smem, as_stream(stream)>>>(heads, rows, ld, x_off, t_off, n_types, w_off, K, X, W, inst, type, T_out, rows_per_patch, patch_stride, patch_offset);
  } else {
    if (smem > 48 * 1024) CPFN_CUDA_TRY(cudaFuncSetAttribute(spfn_post_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    spfn_post_kernel<64><<<grid, kGlueThreads, smem, as_stream(stream)>>>(heads, rows, ld, x_off, t_off, n_types, w_off, K, X, W, inst, type, T_out, rows_per_patch, patch_stride, patch_offset);
  }
  return check_launch();
}

extern "C" int cpfn_spfn_post(const float *heads, long long rows, int ld, int x_off, int t_off, int n_types,
                              int w_off, int K, float *X, float *W, int32_t *inst, int32_t *type,
                              cpfn_stream_t stream) {
  return cpfn_spfn_post_scatter(heads, rows, ld, x_off, t_off, n_types, w_off, K, X, W, nullptr, inst, type, 0, 0, 0, stream);
}
