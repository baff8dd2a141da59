// Ball query for sm_100a.
//
// Semantics (bit-exact with PointNet2/pointnet2_ops/cuda_ops/src/
// ball_query_gpu.cu:9-44, zero-init from src/ball_query.cpp:19-21; SURVEY.md
// appendix A.2): r2 = radius*radius in fp32; scan k ascending; a point is a hit
// iff fma(dz,dz,fma(dx,dx,dy*dy)) < r2; output = first `nsample` hits in index
// order, remaining slots padded with the first hit; no hit -> zeros.
//
// Design: the reference gives one THREAD a query and scans the cloud serially
// with 12-byte strided loads.  Here one WARP owns a query: the cloud is staged
// once per CTA into shared memory as SoA (conflict-free LDS), the 32 lanes test
// 32 consecutive points per step, a ballot + popc prefix hands out output slots
// in ascending index order, and the warp leaves the scan as soon as `nsample`
// hits are found.  Clouds larger than one tile are processed tile by tile with
// the per-query hit count carried in shared memory.
#include <math.h>

#include "common.cuh"

namespace cpfn {
namespace {

constexpr int kBqThreads = 512;
constexpr int kBqWarps = kBqThreads / 32;
constexpr int kBqTile = 8192;     // points per shared-memory tile (96 KB SoA)
constexpr int kBqMaxQpb = 16 * 16;  // queries per block upper bound (state arrays)

__global__ void __launch_bounds__(kBqThreads, 2)
ball_query_kernel(const float *__restrict__ new_xyz, const float *__restrict__ xyz, int N,
                  int S, float radius, int nsample, int qpb, int tile,
                  int32_t *__restrict__ idx) {
  extern __shared__ float s_xyz[];
  __shared__ int s_cnt[kBqMaxQpb];
  __shared__ int s_first[kBqMaxQpb];
  const int b = blockIdx.y;
  const int q0 = blockIdx.x * qpb;
  const int nq = min(qpb, S - q0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float r2 = __fmul_rn(radius, radius);
  const float *p = xyz + static_cast<size_t>(b) * N * 3;
  const float *q = new_xyz + (static_cast<size_t>(b) * S + q0) * 3;
  int32_t *out = idx + (static_cast<size_t>(b) * S + q0) * nsample;
  float *sx = s_xyz, *sy = s_xyz + tile, *sz = s_xyz + 2 * tile;

  for (int i = threadIdx.x; i < nq; i += kBqThreads) { s_cnt[i] = 0; s_first[i] = -1; }
  const unsigned lt_mask = (1u << lane) - 1u;

  for (int t0 = 0; t0 < N; t0 += tile) {
    const int tn = min(tile, N - t0);
    __syncthreads();  // previous tile fully consumed, state arrays visible
    for (int f = threadIdx.x; f < 3 * tn; f += kBqThreads) {
      const int k = f / 3, c = f - 3 * k;
      s_xyz[c * tile + k] = __ldg(p + static_cast<size_t>(t0) * 3 + f);
    }
    __syncthreads();
    int pending = 0;
    for (int qi = warp; qi < nq; qi += kBqWarps) {
      int cnt = s_cnt[qi];
      if (cnt >= nsample) continue;
      int first = s_first[qi];
      const float qx = __ldg(q + 3 * qi), qy = __ldg(q + 3 * qi + 1), qz = __ldg(q + 3 * qi + 2);
      int32_t *o = out + static_cast<size_t>(qi) * nsample;
      for (int base = 0; base < tn && cnt < nsample; base += 128) {
        bool hit[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int k = base + u * 32 + lane;
          const int kc = min(k, tn - 1);
          hit[u] = (k < tn) && (sqdist3(qx, qy, qz, sx[kc], sy[kc], sz[kc]) < r2);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const unsigned bal = __ballot_sync(0xffffffffu, hit[u]);
          if (bal && cnt < nsample) {
            if (cnt == 0) first = t0 + base + u * 32 + __ffs(bal) - 1;
            const int slot = cnt + __popc(bal & lt_mask);
            if (hit[u] && slot < nsample) o[slot] = t0 + base + u * 32 + lane;
            cnt += __popc(bal);
          }
        }
      }
      if (lane == 0) { s_cnt[qi] = cnt; s_first[qi] = first; }
      if (cnt < nsample) pending = 1;
    }
    if (t0 + tile < N) {  // more tiles: stop early when every query is full
      if (!__syncthreads_or(pending)) break;
    }
  }
  __syncthreads();
  // Padding (ball_query_gpu.cu:34-38 fills every slot with the first hit).
  for (int qi = warp; qi < nq; qi += kBqWarps) {
    const int cnt = min(s_cnt[qi], nsample);
    const int first = s_first[qi];
    int32_t *o = out + static_cast<size_t>(qi) * nsample;
    for (int l = cnt + lane; l < nsample; l += 32) o[l] = first < 0 ? 0 : first;
  }
}

// ---- uniform-grid variant -----------------------------------------------------------------------
// The brute-force kernel above tests every point of the cloud until it has `nsample` hits.  Here the
// cloud is first binned into a uniform grid with cells no smaller than the radius (one CTA per cloud:
// bounding box, shared-memory histogram, scan, scatter of (x, y, z, index) records into cell order),
// so a query only tests the points of its 3 x 3 x 3 cell neighbourhood (nine contiguous runs, the three
// x-neighbours being adjacent in cell order).  Hits are recorded in a per-warp BITMAP over the point
// indices, which is then read back in ascending order: the first `nsample` set bits are exactly the
// reference's "first nsample hits in index order", whatever order the grid produced them in.
// Distances use the original coordinates and the same rounding sequence, so the result is bit-identical.
constexpr int kGridMaxAxis = 16;                       // <= 16^3 = 4096 cells
constexpr int kGridMaxCells = kGridMaxAxis * kGridMaxAxis * kGridMaxAxis;
constexpr int kGridBuildThreads = 1024;
constexpr int kGridQueryWarps = 16;
constexpr int kGridMaxN = 32768;                       // bitmap: N/32 words per warp

struct GridHeader {                                     // one per cloud, in the workspace
  float minx, miny, minz, inv_cell;
  int nx, ny, nz, pad;
};

__device__ __forceinline__ int grid_coord(float v, float mn, float inv_cell, int n) {
  int c = static_cast<int>(floorf(__fmul_rn(__fsub_rn(v, mn), inv_cell)));
  return c < 0 ? 0 : (c >= n ? n - 1 : c);
}

__global__ void __launch_bounds__(kGridBuildThreads)
bq_grid_build_kernel(const float *__restrict__ xyz, int N, float radius, GridHeader *__restrict__ hdr,
                     int *__restrict__ cell_start, float4 *__restrict__ sorted) {
  __shared__ int s_cnt[kGridMaxCells + 1];
  __shared__ float s_red[6][32];
  __shared__ int s_scan[32];
  __shared__ GridHeader s_h;
  const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const float *p = xyz + static_cast<size_t>(b) * N * 3;
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int k = t; k < N; k += kGridBuildThreads)
#pragma unroll
    for (int c = 0; c < 3; ++c) { const float v = __ldg(p + 3 * k + c); mn[c] = fminf(mn[c], v); mx[c] = fmaxf(mx[c], v); }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
      mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
    }
    if (lane == 0) { s_red[c][warp] = mn[c]; s_red[3 + c][warp] = mx[c]; }
  }
  for (int i = t; i <= kGridMaxCells; i += kGridBuildThreads) s_cnt[i] = 0;
  __syncthreads();
  if (t == 0) {
    float lo[3], hi[3];
    for (int c = 0; c < 3; ++c) {
      lo[c] = s_red[c][0]; hi[c] = s_red[3 + c][0];
      for (int w = 1; w < kGridBuildThreads / 32; ++w) { lo[c] = fminf(lo[c], s_red[c][w]); hi[c] = fmaxf(hi[c], s_red[3 + c][w]); }
    }
    const float ext = fmaxf(fmaxf(hi[0] - lo[0], hi[1] - lo[1]), hi[2] - lo[2]);
    // cells a little larger than the radius (rounding margin) and at most 16 per axis
    float cell = fmaxf(radius * 1.0001f, ext * (1.0001f / kGridMaxAxis));
    if (!(cell > 0.f) || !isfinite(cell)) cell = 1.f;
    s_h.minx = lo[0]; s_h.miny = lo[1]; s_h.minz = lo[2]; s_h.inv_cell = 1.0f / cell;
    s_h.nx = min(kGridMaxAxis, static_cast<int>((hi[0] - lo[0]) / cell) + 1);
    s_h.ny = min(kGridMaxAxis, static_cast<int>((hi[1] - lo[1]) / cell) + 1);
    s_h.nz = min(kGridMaxAxis, static_cast<int>((hi[2] - lo[2]) / cell) + 1);
    s_h.pad = 0;
    hdr[b] = s_h;
  }
  __syncthreads();
  const GridHeader h = s_h;
  const int ncell = h.nx * h.ny * h.nz;
  for (int k = t; k < N; k += kGridBuildThreads) {
    const float x = __ldg(p + 3 * k), y = __ldg(p + 3 * k + 1), z = __ldg(p + 3 * k + 2);
    const int c = (grid_coord(z, h.minz, h.inv_cell, h.nz) * h.ny + grid_coord(y, h.miny, h.inv_cell, h.ny)) * h.nx +
                  grid_coord(x, h.minx, h.inv_cell, h.nx);
    atomicAdd(&s_cnt[c], 1);
  }
  __syncthreads();
  // exclusive scan of the cell counts (<= 4096 cells, 4 per thread)
  int v[4], sum = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) { const int c = t * 4 + i; v[i] = c < ncell ? s_cnt[c] : 0; sum += v[i]; }
  int inc = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
  if (lane == 31) s_scan[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = s_scan[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += u; }
    s_scan[lane] = w;
  }
  __syncthreads();
  int base = inc - sum + (warp > 0 ? s_scan[warp - 1] : 0);
  int *cs = cell_start + static_cast<size_t>(b) * (kGridMaxCells + 1);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = t * 4 + i;
    if (c < ncell) { s_cnt[c] = base; cs[c] = base; }
    base += v[i];
  }
  if (t == 0) cs[ncell] = N;
  __syncthreads();
  float4 *out = sorted + static_cast<size_t>(b) * N;
  for (int k = t; k < N; k += kGridBuildThreads) {
    const float x = __ldg(p + 3 * k), y = __ldg(p + 3 * k + 1), z = __ldg(p + 3 * k + 2);
    const int c = (grid_coord(z, h.minz, h.inv_cell, h.nz) * h.ny + grid_coord(y, h.miny, h.inv_cell, h.ny)) * h.nx +
                  grid_coord(x, h.minx, h.inv_cell, h.nx);
    out[atomicAdd(&s_cnt[c], 1)] = make_float4(x, y, z, __int_as_float(k));
  }
}

__global__ void __launch_bounds__(kGridQueryWarps * 32)
bq_grid_query_kernel(const float *__restrict__ new_xyz, int N, int S, int s_begin, int s_end, float radius, int nsample,
                     const GridHeader *__restrict__ hdr, const int *__restrict__ cell_start,
                     const float4 *__restrict__ sorted, int32_t *__restrict__ idx) {
  extern __shared__ unsigned int s_bits[];             // [kGridQueryWarps][words]
  const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int words = (N + 31) >> 5;
  unsigned int *bits = s_bits + warp * words;
  const GridHeader h = hdr[b];
  const int *cs = cell_start + static_cast<size_t>(b) * (kGridMaxCells + 1);
  const float4 *pts = sorted + static_cast<size_t>(b) * N;
  const float r2 = __fmul_rn(radius, radius);
  for (int q = s_begin + blockIdx.x * kGridQueryWarps + warp; q < s_end; q += gridDim.x * kGridQueryWarps) {
    const float *c = new_xyz + (static_cast<size_t>(b) * S + q) * 3;
    const float qx = __ldg(c), qy = __ldg(c + 1), qz = __ldg(c + 2);
    for (int i = lane; i < words; i += 32) bits[i] = 0u;
    const int cx = grid_coord(qx, h.minx, h.inv_cell, h.nx), cy = grid_coord(qy, h.miny, h.inv_cell, h.ny),
              cz = grid_coord(qz, h.minz, h.inv_cell, h.nz);
    // lane i < 9 owns the run of cells (cx-1..cx+1, cy + i%3 - 1, cz + i/3 - 1): the three x-neighbours
    // are contiguous in cell order, so the 27-cell neighbourhood is nine contiguous runs of records
    int beg = 0, end = 0;                              // fetched by nine lanes at once: one L2 round trip per query
    if (lane < 9) {
      const int y = cy + lane % 3 - 1, z = cz + lane / 3 - 1;
      if (y >= 0 && y < h.ny && z >= 0 && z < h.nz) {
        const int row = (z * h.ny + y) * h.nx;
        beg = __ldg(cs + row + max(cx - 1, 0));
        end = __ldg(cs + row + min(cx + 1, h.nx - 1) + 1);
      }
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      const int e = __shfl_sync(0xffffffffu, end, i);
      for (int k = __shfl_sync(0xffffffffu, beg, i) + lane; k < e; k += 128) {
        float4 p[4];                                   // four independent record loads in flight per lane
#pragma unroll
        for (int u = 0; u < 4; ++u) p[u] = __ldg(pts + min(k + 32 * u, e - 1));
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (k + 32 * u < e && sqdist3(qx, qy, qz, p[u].x, p[u].y, p[u].z) < r2) {
            const int id = __float_as_int(p[u].w);
            atomicOr(&bits[id >> 5], 1u << (id & 31));
          }
      }
    }
    __syncwarp();
    // read the bitmap back in ascending index order: first `nsample` set bits, padded with the first
    int32_t *o = idx + (static_cast<size_t>(b) * S + q) * nsample;
    int cnt = 0, first = -1;
    for (int w0 = 0; w0 < words && cnt < nsample; w0 += 32) {
      unsigned int word = (w0 + lane < words) ? bits[w0 + lane] : 0u;
      const int n = __popc(word);
      int pre = n;
#pragma unroll
      for (int s2 = 1; s2 < 32; s2 <<= 1) { const int u = __shfl_up_sync(0xffffffffu, pre, s2); if (lane >= s2) pre += u; }
      const int total = __shfl_sync(0xffffffffu, pre, 31);
      int pos = cnt + pre - n;
      if (first < 0 && total > 0) {
        const unsigned int have = __ballot_sync(0xffffffffu, n > 0);
        const int src = __ffs(have) - 1;
        const unsigned int fw = __shfl_sync(0xffffffffu, word, src);
        first = ((w0 + src) << 5) + __ffs(fw) - 1;
      }
      while (word && pos < nsample) {
        const int bit = __ffs(word) - 1;
        o[pos++] = ((w0 + lane) << 5) + bit;
        word &= word - 1;
      }
      cnt += total;
    }
    if (cnt > nsample) cnt = nsample;
    for (int i = cnt + lane; i < nsample; i += 32) o[i] = first < 0 ? 0 : first;
    __syncwarp();
  }
}

}  // namespace
}  // namespace cpfn

extern "C" int cpfn_ball_query(const float *new_xyz, const float *xyz, int B, int N, int S,
                               float radius, int nsample, int32_t *idx, cpfn_stream_t stream) {
  using namespace cpfn;
  if (B < 0 || N < 0 || S < 0 || nsample < 0) return CPFN_EINVAL;
  if (B == 0 || S == 0 || nsample == 0) return CPFN_OK;
  if (!new_xyz || !idx || (N > 0 && !xyz)) return CPFN_EINVAL;
  if (B > 65535) return CPFN_EINVAL;
  cudaStream_t st = as_stream(stream);
  if (N == 0) {
    CPFN_CUDA_TRY(cudaMemsetAsync(idx, 0, sizeof(int32_t) * size_t(B) * S * nsample, st));
    return CPFN_OK;
  }
  // Queries per block: enough blocks to cover the SMs about twice, one to
  // sixteen queries per warp.
  const int sms = sm_count() > 0 ? sm_count() : 148;
  int qpw = 1;
  while (qpw < 16 && static_cast<long long>(B) * ((S + kBqWarps * qpw - 1) / (kBqWarps * qpw)) >
                         2LL * sms)
    qpw *= 2;
  const int qpb = kBqWarps * qpw;
  const int tile = N < kBqTile ? ((N + 31) & ~31) : kBqTile;
  const size_t smem = 3u * static_cast<size_t>(tile) * sizeof(float);
  if (smem + 4096 > 48 * 1024)
    CPFN_CUDA_TRY(cudaFuncSetAttribute(ball_query_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       3 * kBqTile * static_cast<int>(sizeof(float))));
  dim3 grid((S + qpb - 1) / qpb, B);
  ball_query_kernel<<<grid, kBqThreads, smem, st>>>(new_xyz, xyz, N, S, radius, nsample, qpb,
                                                     tile, idx);
  return check_launch();
}

// Uniform-grid ball query: same result as cpfn_ball_query, far fewer distance tests.  Workspace:
// per cloud a header, the cell table and the cell-sorted (x, y, z, index) records.
extern "C" size_t cpfn_ball_query_grid_workspace_bytes(int B, int N) {
  using namespace cpfn;
  if (B <= 0 || N <= 0) return 0;
  return static_cast<size_t>(B) * (sizeof(GridHeader) + sizeof(int) * (kGridMaxCells + 1) + 16) +
         static_cast<size_t>(B) * N * sizeof(float4) + 256;
}

namespace cpfn {
namespace {
struct GridWs { float4 *sorted; GridHeader *hdr; int *cell_start; };
GridWs grid_ws(void *workspace, int B, int N) {
  unsigned char *w = static_cast<unsigned char *>(workspace);
  GridWs g;
  g.sorted = reinterpret_cast<float4 *>(w);
  w += static_cast<size_t>(B) * N * sizeof(float4);
  g.hdr = reinterpret_cast<GridHeader *>(w);
  w += static_cast<size_t>(B) * sizeof(GridHeader);
  g.cell_start = reinterpret_cast<int *>(w);
  return g;
}
bool grid_applies(int B, int N, float radius) { return N >= 2048 && N <= kGridMaxN && radius > 0.f && B <= 65535; }
}  // namespace
}  // namespace cpfn

// The grid depends on the cloud and the radius only -- not on the queries -- so a caller can build it while
// the queries are still being produced (the SA layer's farthest point sampling) and query it afterwards.
extern "C" int cpfn_ball_query_grid_build(const float *xyz, int B, int N, float radius, void *workspace,
                                          size_t workspace_bytes, cpfn_stream_t stream) {
  using namespace cpfn;
  if (B < 0 || N < 0) return CPFN_EINVAL;
  if (B == 0 || !grid_applies(B, N, radius)) return CPFN_OK;                 // the query will use the scan kernel
  if (!xyz) return CPFN_EINVAL;
  if (!workspace || workspace_bytes < cpfn_ball_query_grid_workspace_bytes(B, N)) return CPFN_EWORKSPACE;
  const GridWs g = grid_ws(workspace, B, N);
  bq_grid_build_kernel<<<B, kGridBuildThreads, 0, as_stream(stream)>>>(xyz, N, radius, g.hdr, g.cell_start, g.sorted);
  return check_launch();
}

extern "C" int cpfn_ball_query_grid_query_range(const float *new_xyz, const float *xyz, int B, int N, int S,
                                                int s_begin, int s_count, float radius, int nsample, int32_t *idx,
                                                void *workspace, size_t workspace_bytes, cpfn_stream_t stream) {
  using namespace cpfn;
  if (B < 0 || N < 0 || S < 0 || nsample < 0 || s_begin < 0 || s_count < 0 || s_begin + s_count > S) return CPFN_EINVAL;
  if (B == 0 || s_count == 0 || nsample == 0) return CPFN_OK;
  if (!grid_applies(B, N, radius)) return CPFN_EINVAL;                       // ranges exist for the grid kernel only
  if (!new_xyz || !idx) return CPFN_EINVAL;
  if (!workspace || workspace_bytes < cpfn_ball_query_grid_workspace_bytes(B, N)) return CPFN_EWORKSPACE;
  cudaStream_t st = as_stream(stream);
  const GridWs g = grid_ws(workspace, B, N);
  const size_t smem = sizeof(unsigned int) * kGridQueryWarps * static_cast<size_t>((N + 31) >> 5);
  if (smem > 48 * 1024)
    CPFN_CUDA_TRY(cudaFuncSetAttribute(bq_grid_query_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem)));
  const int sms = sm_count() > 0 ? sm_count() : 148;
  int gx = (s_count + kGridQueryWarps - 1) / kGridQueryWarps;
  const int cap = (4 * sms + B - 1) / B;
  if (gx > cap) gx = cap < 1 ? 1 : cap;
  bq_grid_query_kernel<<<dim3(gx, B), kGridQueryWarps * 32, smem, st>>>(new_xyz, N, S, s_begin, s_begin + s_count, radius,
                                                                        nsample, g.hdr, g.cell_start, g.sorted, idx);
  return check_launch();
}

extern "C" int cpfn_ball_query_grid_query(const float *new_xyz, const float *xyz, int B, int N, int S, float radius,
                                          int nsample, int32_t *idx, void *workspace, size_t workspace_bytes,
                                          cpfn_stream_t stream) {
  using namespace cpfn;
  if (B < 0 || N < 0 || S < 0 || nsample < 0) return CPFN_EINVAL;
  if (B == 0 || S == 0 || nsample == 0) return CPFN_OK;
  if (!grid_applies(B, N, radius))                                           // small / huge clouds: the scan kernel
    return cpfn_ball_query(new_xyz, xyz, B, N, S, radius, nsample, idx, stream);
  return cpfn_ball_query_grid_query_range(new_xyz, xyz, B, N, S, 0, S, radius, nsample, idx, workspace, workspace_bytes,
                                          stream);
}

extern "C" int cpfn_ball_query_grid(const float *new_xyz, const float *xyz, int B, int N, int S, float radius,
                                    int nsample, int32_t *idx, void *workspace, size_t workspace_bytes,
                                    cpfn_stream_t stream) {
  if (B < 0 || N < 0 || S < 0 || nsample < 0) return CPFN_EINVAL;
  if (B == 0 || S == 0 || nsample == 0) return CPFN_OK;
  const int rc = cpfn_ball_query_grid_build(xyz, B, N, radius, workspace, workspace_bytes, stream);
  if (rc != CPFN_OK) return rc;
  return cpfn_ball_query_grid_query(new_xyz, xyz, B, N, S, radius, nsample, idx, workspace, workspace_bytes, stream);
}
