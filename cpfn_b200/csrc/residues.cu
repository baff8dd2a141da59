// Point-to-primitive residues (SURVEY 8a row a14 and the residue part of row f3).
//
//   cpfn_primitive_residues : compute_residue_loss (SPFN/losses_implementation.py:351-387) -- for every matched
//       primitive and every one of its points the squared distance to the plane / sphere / cylinder / cone
//       (plane_fitter.py:54-55, sphere_fitter.py:61-62, cylinder_fitter.py:85-89, cone_fitter.py:98-103), all
//       requested types in ONE pass over the points, plus the per-primitive means.
//   cpfn_p_coverage         : compute_P_coverage (SPFN/metric_implementation.py:409-415) -- per point the
//       smallest residual sqrt_safe(residue of the primitive's own type) over all primitives, compared with the
//       thresholds and averaged; the reference expands P to [B,K,N,3] and materialises [B,K,N,T] for it.
//
// The arithmetic follows torch's element-wise kernels step by step (separately rounded products, sums over the
// 3 coordinates left to right, sqrt_safe(x) = sqrt(|x| + 1e-10), acos_safe clamp +-(1 - 1e-6), F.normalize with
// eps 1e-12); acosf / sinf are the CUDA math library's, as in torch's CUDA kernels.
#include <math.h>

#include "common.cuh"

namespace cpfn {
namespace {

constexpr int kResThreads = 256;

struct Prim {                       // parameters of one matched primitive slot
  float pn[3], pc;                  // plane normal, offset
  float sc[3], sr2;                 // sphere centre, radius^2
  float ya[3], yc[3], yr2;          // cylinder axis, centre, radius^2
  float ka[3], kx[3], kh;           // cone apex, axis, half angle
};

__device__ __forceinline__ void load3(const float *p, size_t i, float *o) {
  if (p) { o[0] = __ldg(p + 3 * i); o[1] = __ldg(p + 3 * i + 1); o[2] = __ldg(p + 3 * i + 2); } else { o[0] = o[1] = o[2] = 0.f; }
}

__device__ __forceinline__ Prim load_prim(const cpfn_primitive_params_t &q, size_t i) {
  Prim r;
  load3(q.plane_normal, i, r.pn);   r.pc = q.plane_center ? __ldg(q.plane_center + i) : 0.f;
  load3(q.sphere_center, i, r.sc);  r.sr2 = q.sphere_radius_squared ? __ldg(q.sphere_radius_squared + i) : 0.f;
  load3(q.cylinder_axis, i, r.ya);  load3(q.cylinder_center, i, r.yc);
  r.yr2 = q.cylinder_radius_squared ? __ldg(q.cylinder_radius_squared + i) : 0.f;
  load3(q.cone_apex, i, r.ka);      load3(q.cone_axis, i, r.kx);
  r.kh = q.cone_half_angle ? __ldg(q.cone_half_angle + i) : 0.f;
  return r;
}

__device__ __forceinline__ float sum3(float a, float b, float c) { return __fadd_rn(__fadd_rn(a, b), c); }
__device__ __forceinline__ float dot3(const float *a, float x, float y, float z) {
  return sum3(__fmul_rn(x, a[0]), __fmul_rn(y, a[1]), __fmul_rn(z, a[2]));
}
__device__ __forceinline__ float sqrt_safe(float x) { return __fsqrt_rn(__fadd_rn(fabsf(x), 1e-10f)); }
__device__ __forceinline__ float sq(float x) { return __fmul_rn(x, x); }

// class ids: 0 plane, 1 sphere, 2 cylinder, 3 cone
__device__ __forceinline__ float residue(const Prim &r, int cls, float x, float y, float z) {
  if (cls == 0) return sq(__fsub_rn(dot3(r.pn, x, y, z), r.pc));                       // (sum(p*n) - c)^2
  if (cls == 1) {
    const float dx = __fsub_rn(x, r.sc[0]), dy = __fsub_rn(y, r.sc[1]), dz = __fsub_rn(z, r.sc[2]);
    return sq(__fsub_rn(sqrt_safe(sum3(sq(dx), sq(dy), sq(dz))), sqrt_safe(r.sr2)));
  }
  if (cls == 2) {
    const float dx = __fsub_rn(x, r.yc[0]), dy = __fsub_rn(y, r.yc[1]), dz = __fsub_rn(z, r.yc[2]);
    const float d2 = sum3(sq(dx), sq(dy), sq(dz));
    const float dn = dot3(r.ya, dx, dy, dz);
    return sq(__fsub_rn(sqrt_safe(__fsub_rn(d2, sq(dn))), sqrt_safe(r.yr2)));
  }
  const float vx = __fsub_rn(x, r.ka[0]), vy = __fsub_rn(y, r.ka[1]), vz = __fsub_rn(z, r.ka[2]);
  const float v2 = sum3(sq(vx), sq(vy), sq(vz));
  const float inv = fmaxf(__fsqrt_rn(v2), 1e-12f);                                      // F.normalize(eps=1e-12)
  float c = dot3(r.kx, __fdiv_rn(vx, inv), __fdiv_rn(vy, inv), __fdiv_rn(vz, inv));
  c = fminf(fmaxf(c, -1.0f + 1e-6f), 1.0f - 1e-6f);                                     // acos_safe
  const float alpha = acosf(c);
  const float s = sinf(fminf(fabsf(__fsub_rn(alpha, r.kh)), 1.57079632679489661923f));
  return __fmul_rn(sq(s), v2);
}

// grid (chunks of points, K, B).  per_point [B,K,n_pts,T] and/or mean [B,K,T].
__global__ void __launch_bounds__(kResThreads)
residues_kernel(cpfn_primitive_params_t q, const int32_t *__restrict__ match, const float *__restrict__ points,
                long long stride_b, long long stride_k, int Kp, int K, int n_pts, int T, int4 classes,
                float *__restrict__ per_point, float *__restrict__ mean_acc) {
  const int b = blockIdx.z, k = blockIdx.y;
  const int slot = __ldg(match + static_cast<size_t>(b) * K + k);
  const Prim r = load_prim(q, static_cast<size_t>(b) * Kp + slot);
  const int cls[4] = {classes.x, classes.y, classes.z, classes.w};
  const float *pts = points + b * stride_b + k * stride_k;
  float part[4] = {0.f, 0.f, 0.f, 0.f};
  for (int i = blockIdx.x * kResThreads + threadIdx.x; i < n_pts; i += gridDim.x * kResThreads) {
    const float x = __ldg(pts + 3ll * i), y = __ldg(pts + 3ll * i + 1), z = __ldg(pts + 3ll * i + 2);
#pragma unroll
    for (int t = 0; t < 4; ++t)
      if (t < T) {
        const float v = residue(r, cls[t], x, y, z);
        if (per_point) per_point[((static_cast<size_t>(b) * K + k) * n_pts + i) * T + t] = v;
        part[t] += v;
      }
  }
  if (!mean_acc) return;
  __shared__ float s_part[4][kResThreads / 32];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    float v = part[t];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s_part[t][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < T) {
    float v = 0.f;
    for (int w = 0; w < kResThreads / 32; ++w) v += s_part[threadIdx.x][w];
    atomicAdd(mean_acc + (static_cast<size_t>(b) * K + k) * T + threadIdx.x, v / static_cast<float>(n_pts));
  }
}

// grid (chunks of points, B): min over the K primitives of sqrt_safe(residue of the primitive's own class),
// counted against up to 4 thresholds.  count [B, n_eps] (float, number of covered points).
__global__ void __launch_bounds__(kResThreads)
p_coverage_kernel(cpfn_primitive_params_t q, const int32_t *__restrict__ match, const int32_t *__restrict__ prim_class,
                  const float *__restrict__ P, int Kp, int K, int N, int n_eps, float4 eps, float *__restrict__ count) {
  extern __shared__ unsigned char s_raw[];
  Prim *s_prim = reinterpret_cast<Prim *>(s_raw);
  int *s_cls = reinterpret_cast<int *>(s_prim + K);
  const int b = blockIdx.y;
  for (int k = threadIdx.x; k < K; k += kResThreads) {
    s_prim[k] = load_prim(q, static_cast<size_t>(b) * Kp + __ldg(match + static_cast<size_t>(b) * K + k));
    s_cls[k] = __ldg(prim_class + static_cast<size_t>(b) * K + k);
  }
  __syncthreads();
  const float e[4] = {eps.x, eps.y, eps.z, eps.w};
  int hit[4] = {0, 0, 0, 0};
  for (int i = blockIdx.x * kResThreads + threadIdx.x; i < N; i += gridDim.x * kResThreads) {
    const float *p = P + (static_cast<size_t>(b) * N + i) * 3;
    const float x = __ldg(p), y = __ldg(p + 1), z = __ldg(p + 2);
    float best = INFINITY;
    for (int k = 0; k < K; ++k) best = fminf(best, sqrt_safe(residue(s_prim[k], s_cls[k], x, y, z)));
#pragma unroll
    for (int t = 0; t < 4; ++t) hit[t] += (t < n_eps && best < e[t]) ? 1 : 0;
  }
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    int v = hit[t];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (t < n_eps && (threadIdx.x & 31) == 0 && v) atomicAdd(count + static_cast<size_t>(b) * n_eps + t, static_cast<float>(v));
  }
}

bool classes_ok(const int *class_ids, int T) {
  if (!class_ids || T < 1 || T > 4) return false;
  for (int t = 0; t < T; ++t)
    if (class_ids[t] < 0 || class_ids[t] > 3) return false;
  return true;
}

}  // namespace
}  // namespace cpfn

using namespace cpfn;

extern "C" int cpfn_primitive_residues(const cpfn_primitive_params_t *params, const int32_t *matching, const float *points,
                                       long long stride_b, long long stride_k, int B, int Kp, int K, int n_pts,
                                       const int *class_ids, int T, float *per_point, float *mean, cpfn_stream_t stream) {
  if (!params || !matching || !points || B <= 0 || Kp <= 0 || K <= 0 || K > 65535 || B > 65535 || n_pts <= 0 ||
      !classes_ok(class_ids, T) || (!per_point && !mean))
    return CPFN_EINVAL;
  cudaStream_t s = as_stream(stream);
  if (mean) CPFN_CUDA_TRY(cudaMemsetAsync(mean, 0, sizeof(float) * static_cast<size_t>(B) * K * T, s));
  int chunks = (n_pts + 4 * kResThreads - 1) / (4 * kResThreads);
  if (chunks > 64) chunks = 64;
  const int4 cls = make_int4(class_ids[0], T > 1 ? class_ids[1] : 0, T > 2 ? class_ids[2] : 0, T > 3 ? class_ids[3] : 0);
  residues_kernel<<<dim3(chunks, K, B), kResThreads, 0, s>>>(*params, matching, points, stride_b, stride_k, Kp, K, n_pts, T,
                                                            cls, per_point, mean);
  return check_launch();
}

extern "C" int cpfn_p_coverage(const cpfn_primitive_params_t *params, const int32_t *matching, const int32_t *prim_class,
                               const float *P, int B, int Kp, int K, int N, const float *epsilons_host, int n_eps,
                               float *count, cpfn_stream_t stream) {
  if (!params || !matching || !prim_class || !P || !count || !epsilons_host || B <= 0 || B > 65535 || Kp <= 0 || K <= 0 ||
      K > 512 || N <= 0 || n_eps < 1 || n_eps > 4)
    return CPFN_EINVAL;
  cudaStream_t s = as_stream(stream);
  CPFN_CUDA_TRY(cudaMemsetAsync(count, 0, sizeof(float) * static_cast<size_t>(B) * n_eps, s));
  const int sms = sm_count();
  if (sms <= 0) return CPFN_ELAUNCH;
  int chunks = (N + kResThreads - 1) / kResThreads;
  const int cap = (4 * sms + B - 1) / B;
  if (chunks > cap) chunks = cap;
  const float4 eps = make_float4(epsilons_host[0], n_eps > 1 ? epsilons_host[1] : 0.f, n_eps > 2 ? epsilons_host[2] : 0.f,
                                 n_eps > 3 ? epsilons_host[3] : 0.f);
  const size_t smem = (sizeof(Prim) + sizeof(int)) * K;
  p_coverage_kernel<<<dim3(chunks, B), kResThreads, smem, s>>>(*params, matching, prim_class, P, Kp, K, N, n_eps, eps, count);
  return check_launch();
}
