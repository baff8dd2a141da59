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
