// SPFN weighted total-least-squares primitive fitters for sm_100a: plane, sphere,
// cylinder and cone parameters of every (cloud, instance slot) pair in ONE call.
//
// Semantics follow the reference fitters (paths relative to the reference tree):
//   plane     SPFN/plane_fitter.py:9-17    -> SPFN/geometry_utils.py:74-84 (weighted_plane_fitting)
//   sphere    SPFN/sphere_fitter.py:9-19   -> SPFN/geometry_utils.py:209-223, 121-142
//   cylinder  SPFN/cylinder_fitter.py:10-28 (TLS on the normals, consistent frame
//             SPFN/geometry_utils.py:8-27, 2-D circle fit)
//   cone      SPFN/cone_fitter.py:12-36    (guarded LS apex, plane fit of the normals, half angle)
//   TLS       SPFN/differentiable_tls.py:200-209 (last right-singular vector of sum w a a^T)
// with their constants: division eps 1e-10, LS row weights
// sqrt(max(w,1e-10)), condition-number cap 1e5, ridge 1e-8, acos clamp 1-1e-6, half-angle
// clamp [1e-3, pi/2-1e-3].
//
// Design (not the reference's).  The reference tiles P and X K times ([B*K,N,3], 44 MB each
// at B=16,K=28), materialises [B*K,N,3,3] outer products (132 MB, twice) and calls batched
// SVD / solve.  Here nothing is tiled or materialised: three streaming passes over the
// membership matrix W [B,N,K] (the only large operand) accumulate per-(b,k) weighted moments,
// and three tiny solve kernels do the 3x3 / 2x2 eigen problems and linear solves in registers
// (fp64 Jacobi).
//   pass 1: sum w, sum w p, sum w |p|^2, sum w x, sum w x x^T, sum w' x x^T, sum w' x (p.x)
//   solve1: means (rounded to fp32 like the reference's), cylinder axis, cone apex, cone axis (the scatter of
//           the normals about their fp32-rounded mean follows from the raw normal moments by exact algebra in
//           fp64: the normals are unit vectors, so the shift costs no accuracy worth the name)
//   pass 2: centred moments about the per-slot means: S = sum w d d^T, S' = sum w' d d^T,
//           sum w' d, third-order T' = sum w' d d d; and, with the apex and the axis of solve1, the cone's
//           sum w (dir.axis), sum w acos|dir.axis|
//   solve2: plane, sphere, cylinder (the 2-D circle fit is obtained by projecting the 3-D
//           centred moments onto the cylinder frame -- no pass over the points is needed
//           once the axis is known), cone sign fix and half angle.
// Two passes over W -- SURVEY 8(d)'s model of 2 B N (4K + 24) bytes; round 1 took a third pass for the cone.
// (w' = max(w, 1e-10), d = p - mean.)  Thread t of a CTA owns slot k = t % K and point lane
// g = t / K, so a warp reads W as one contiguous stream (coalesced, every byte used once);
// point coordinates are staged once per CTA in shared memory as float4.  Each thread sums at
// most 32 points in fp32 (weights preloaded: 32 independent loads in flight); cross-thread and cross-CTA
// combination is fp64 and deterministic (per-CTA partials, no atomics).
#include <math.h>

#include "common.cuh"

namespace cpfn {
namespace {

constexpr int kTlsThreads = 256;
constexpr int kMaxCP = 1024;      // points staged per sub-chunk
constexpr int kPPT = 32;          // points per thread and sub-chunk (weights preloaded in registers)
constexpr int kF1 = 23, kF2 = 27, kF3 = 2, kF4 = 32;   // kF4: raw moments up to third order (training path)
constexpr int kFP = 32;           // padded feature stride of the partial arrays
constexpr int kState = 24;        // doubles per (b,k)
// state layout
constexpr int ST_SW = 0, ST_DENOM = 1, ST_MU = 2, ST_M2 = 5, ST_MUX = 6, ST_CYLN = 9, ST_APEX = 12,
              ST_S1 = 15, ST_CONEAX = 18;

struct TlsGeom {
  int G, ppt, CP, chunks, iters;
};

__host__ __device__ inline int imin(int a, int b) { return a < b ? a : b; }
__host__ __device__ inline int imax(int a, int b) { return a > b ? a : b; }

TlsGeom tls_geom(int B, int N, int K, int sms) {
  TlsGeom g;
  g.G = imax(1, kTlsThreads / K);
  long long want = (static_cast<long long>(B) * N + 2LL * sms * g.G - 1) / (2LL * sms * g.G);
  g.ppt = static_cast<int>(want < 8 ? 8 : (want > kPPT ? kPPT : want));
  g.ppt = imin((g.ppt + 7) & ~7, imax(8, (kMaxCP / g.G) & ~7));   // a multiple of 8 (uniform 8-point blocks)
  g.CP = g.G * g.ppt;
  const int SC = (N + g.CP - 1) / g.CP;
  const int cap = imax(1, (4 * sms + B - 1) / B);
  g.chunks = imin(SC, cap);
  g.iters = (SC + g.chunks - 1) / g.chunks;
  g.chunks = (SC + g.iters - 1) / g.iters;
  return g;
}

template <int PASS> struct NFeat { static constexpr int F = PASS == 1 ? kF1 : (PASS == 2 ? kF2 : (PASS == 3 ? kF3 : kF4)); };

template <int PASS>
__global__ void __launch_bounds__(kTlsThreads, 2)
tls_pass_kernel(const float *__restrict__ P, const float *__restrict__ X,
                const float *__restrict__ W, const double *__restrict__ state,
                double *__restrict__ part, int N, int K, int G, int CP, int iters, int chunks) {
  constexpr int F = NFeat<PASS>::F;
  __shared__ float s_piv[4];                 // PASS 1: the CTA's pivot for the first and second normal moments
  extern __shared__ float4 s_pts[];          // [CP] positions (+ |p|^2), [CP] normals (+ p.x)
  float4 *sp = s_pts, *sx = s_pts + CP;
  double *red = reinterpret_cast<double *>(s_pts + 2 * CP);   // [kTlsThreads][8]
  double *tot = red + kTlsThreads * 8;                        // [K][kFP] running fp64 totals
  const int b = blockIdx.y, chunk = blockIdx.x, t = threadIdx.x;
  const bool active = t < G * K;
  const int k = active ? t % K : 0, g = t / K;
  const float *Pb = P + static_cast<size_t>(b) * N * 3;
  const float *Xb = X ? X + static_cast<size_t>(b) * N * 3 : nullptr;
  const float *Wb = W + static_cast<size_t>(b) * N * K;
  const int ppt = CP / G;

  float c0 = 0.f, c1 = 0.f, c2 = 0.f, e0 = 0.f, e1 = 0.f, e2 = 0.f, a0 = 0.f, a1 = 0.f, a2 = 0.f;   // per-slot constants
  if (PASS == 2) {
    const double *st = state + (static_cast<size_t>(b) * K + k) * kState;
    c0 = static_cast<float>(st[ST_MU]); c1 = static_cast<float>(st[ST_MU + 1]); c2 = static_cast<float>(st[ST_MU + 2]);
    a0 = static_cast<float>(st[ST_APEX]); a1 = static_cast<float>(st[ST_APEX + 1]); a2 = static_cast<float>(st[ST_APEX + 2]);
    e0 = static_cast<float>(st[ST_CONEAX]); e1 = static_cast<float>(st[ST_CONEAX + 1]); e2 = static_cast<float>(st[ST_CONEAX + 2]);
  }
  for (int o = t; o < K * kFP; o += kTlsThreads) tot[o] = 0.0;

  for (int it = 0; it < iters; ++it) {
    const int n0 = (chunk * iters + it) * CP;
    if (n0 >= N) break;
    const int cn = imin(CP, N - n0);
    // All of this thread's membership weights first: up to 32 independent loads in flight.
    float wv[kPPT];
    {
      const float *wp = Wb + static_cast<size_t>(n0) * K + k;
#pragma unroll
      for (int i = 0; i < kPPT; ++i) {
        const int j = g + i * G;
        wv[i] = (active && i < ppt && j < cn) ? __ldg(wp + static_cast<size_t>(j) * K) : 0.f;
      }
    }
    __syncthreads();
    for (int i = t; i < CP; i += kTlsThreads) {          // entries past the end of the cloud are zeros
      float px = 0.f, py = 0.f, pz = 0.f, xx = 0.f, xy = 0.f, xz = 0.f;
      if (i < cn) {
        const float *p = Pb + static_cast<size_t>(n0 + i) * 3;
        px = __ldg(p); py = __ldg(p + 1); pz = __ldg(p + 2);
        if (Xb) {
          const float *x = Xb + static_cast<size_t>(n0 + i) * 3;
          xx = __ldg(x); xy = __ldg(x + 1); xz = __ldg(x + 2);
        }
      }
      sp[i] = make_float4(px, py, pz, px * px + py * py + pz * pz);
      sx[i] = make_float4(xx, xy, xz, px * xx + py * xy + pz * xz);
    }
    __syncthreads();
    if (PASS == 1 && it == 0) {
      // Pivot of this CTA's normal moments: the plain mean of the normals it staged first.  sum w x and sum w x x^T
      // are accumulated about it and shifted back to the origin in fp64 when the CTA writes its partials: when the
      // normals a slot weights are concentrated around the cloud's mean normal -- near-uniform memberships -- the
      // fp32 terms are small and the scatter about the slot mean, which the cone axis comes from, keeps its digits.
      float a = 0.f, bq = 0.f, c = 0.f;
      for (int i = t; i < cn; i += kTlsThreads) { const float4 x = sx[i]; a += x.x; bq += x.y; c += x.z; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o); bq += __shfl_xor_sync(0xffffffffu, bq, o); c += __shfl_xor_sync(0xffffffffu, c, o);
      }
      float *wsum = reinterpret_cast<float *>(red);            // [8 warps][3], free at this point
      if ((t & 31) == 0) { wsum[(t >> 5) * 3] = a; wsum[(t >> 5) * 3 + 1] = bq; wsum[(t >> 5) * 3 + 2] = c; }
      __syncthreads();
      if (t < 3) {
        float v = 0.f;
        for (int w8 = 0; w8 < kTlsThreads / 32; ++w8) v += wsum[w8 * 3 + t];
        s_piv[t] = v / static_cast<float>(cn);
      }
      __syncthreads();
    }
    const float pv0 = PASS == 1 ? s_piv[0] : 0.f, pv1 = PASS == 1 ? s_piv[1] : 0.f, pv2 = PASS == 1 ? s_piv[2] : 0.f;
    float acc[F];
#pragma unroll
    for (int f = 0; f < F; ++f) acc[f] = 0.f;
    // this thread's valid points in the sub-chunk: j = g + i*G < cn  <=>  i < lim.  Blocks of 8 points are
    // skipped uniformly (ppt is a multiple of 8); inside a block invalid pairs carry w = w' = 0.
    const int lim = active ? imax(0, (cn - g + G - 1) / G) : 0;
    const float4 *pp = sp + g, *xp = sx + g;
#pragma unroll
    for (int i = 0; i < kPPT; ++i) {
      if ((i & 7) == 0 && i >= ppt) break;
      const bool ok = i < lim;
      const float w = wv[i];
      const float4 p = pp[i * G];
      if (PASS == 1) {
        const float4 x = xp[i * G];
        const float wc = ok ? fmaxf(w, 1e-10f) : 0.f;
        acc[0] += w;
        acc[1] = fmaf(w, p.x, acc[1]); acc[2] = fmaf(w, p.y, acc[2]); acc[3] = fmaf(w, p.z, acc[3]);
        acc[4] = fmaf(w, p.w, acc[4]);
        const float qx = x.x - pv0, qy = x.y - pv1, qz = x.z - pv2;      // about the CTA's pivot
        const float wx = w * qx, wy = w * qy, wz = w * qz;
        acc[5] += wx; acc[6] += wy; acc[7] += wz;
        acc[8] = fmaf(wx, qx, acc[8]); acc[9] = fmaf(wx, qy, acc[9]); acc[10] = fmaf(wx, qz, acc[10]);
        acc[11] = fmaf(wy, qy, acc[11]); acc[12] = fmaf(wy, qz, acc[12]); acc[13] = fmaf(wz, qz, acc[13]);
        const float vx = wc * x.x, vy = wc * x.y, vz = wc * x.z;
        acc[14] = fmaf(vx, x.x, acc[14]); acc[15] = fmaf(vx, x.y, acc[15]); acc[16] = fmaf(vx, x.z, acc[16]);
        acc[17] = fmaf(vy, x.y, acc[17]); acc[18] = fmaf(vy, x.z, acc[18]); acc[19] = fmaf(vz, x.z, acc[19]);
        acc[20] = fmaf(vx, x.w, acc[20]); acc[21] = fmaf(vy, x.w, acc[21]); acc[22] = fmaf(vz, x.w, acc[22]);
      } else if (PASS == 2) {
        const float wc = ok ? fmaxf(w, 1e-10f) : 0.f;
        const float dx = p.x - c0, dy = p.y - c1, dz = p.z - c2;
        const float wx = w * dx, wy = w * dy, wz = w * dz;
        acc[0] = fmaf(wx, dx, acc[0]); acc[1] = fmaf(wx, dy, acc[1]); acc[2] = fmaf(wx, dz, acc[2]);
        acc[3] = fmaf(wy, dy, acc[3]); acc[4] = fmaf(wy, dz, acc[4]); acc[5] = fmaf(wz, dz, acc[5]);
        const float vx = wc * dx, vy = wc * dy, vz = wc * dz;
        acc[12] += vx; acc[13] += vy; acc[14] += vz;
        const float uxx = vx * dx, uxy = vx * dy, uxz = vx * dz, uyy = vy * dy, uyz = vy * dz, uzz = vz * dz;
        acc[6] += uxx; acc[7] += uxy; acc[8] += uxz; acc[9] += uyy; acc[10] += uyz; acc[11] += uzz;
        acc[15] = fmaf(uxx, dx, acc[15]); acc[16] = fmaf(uxx, dy, acc[16]); acc[17] = fmaf(uxx, dz, acc[17]);
        acc[18] = fmaf(uxy, dy, acc[18]); acc[19] = fmaf(uxy, dz, acc[19]); acc[20] = fmaf(uxz, dz, acc[20]);
        acc[21] = fmaf(uyy, dy, acc[21]); acc[22] = fmaf(uyy, dz, acc[22]); acc[23] = fmaf(uyz, dz, acc[23]);
        acc[24] = fmaf(uzz, dz, acc[24]);
        // cone_fitter.py:25-33: dir = normalize(p - apex, eps 1e-12); dot = axis . dir; acos(clamp |dot|)
        const float hx = p.x - a0, hy = p.y - a1, hz = p.z - a2;
        const float n2 = fmaf(hz, hz, fmaf(hy, hy, hx * hx));
        const float inv = n2 > 1e-24f ? rsqrtf(n2) : 1e12f;              // 1 / max(|v|, 1e-12)
        const float dot = fmaf(e2, hz, fmaf(e1, hy, e0 * hx)) * inv;
        acc[25] = fmaf(w, dot, acc[25]);
        const float a = fminf(fabsf(dot), 1.0f - 1e-6f);
        // acos on [0,1): sqrt(1-a) * P7(a) (Abramowitz & Stegun 4.4.46, |error| <= 2e-8): the sum below is
        // divided by sum w afterwards, far inside the 1e-5 tolerance, at a third of acosf's instruction count
        float poly = fmaf(-0.0012624911f, a, 0.0066700901f);
        poly = fmaf(poly, a, -0.0170881256f);
        poly = fmaf(poly, a, 0.0308918810f);
        poly = fmaf(poly, a, -0.0501743046f);
        poly = fmaf(poly, a, 0.0889789874f);
        poly = fmaf(poly, a, -0.2145988016f);
        poly = fmaf(poly, a, 1.5707963050f);
        acc[26] = fmaf(w, sqrtf(1.0f - a) * poly, acc[26]);
      } else if (PASS == 4) {
        // raw moments  sum w * [1, p, p p^T, p p p, x, x x^T, x (p.x)]  (feature order of cpfn_weighted_moments)
        const float4 x = xp[i * G];
        const float wx = w * p.x, wy = w * p.y, wz = w * p.z;
        acc[0] += w; acc[1] += wx; acc[2] += wy; acc[3] += wz;
        const float uxx = wx * p.x, uxy = wx * p.y, uxz = wx * p.z, uyy = wy * p.y, uyz = wy * p.z, uzz = wz * p.z;
        acc[4] += uxx; acc[5] += uxy; acc[6] += uxz; acc[7] += uyy; acc[8] += uyz; acc[9] += uzz;
        acc[10] = fmaf(uxx, p.x, acc[10]); acc[11] = fmaf(uxx, p.y, acc[11]); acc[12] = fmaf(uxx, p.z, acc[12]);
        acc[13] = fmaf(uxy, p.y, acc[13]); acc[14] = fmaf(uxy, p.z, acc[14]); acc[15] = fmaf(uxz, p.z, acc[15]);
        acc[16] = fmaf(uyy, p.y, acc[16]); acc[17] = fmaf(uyy, p.z, acc[17]); acc[18] = fmaf(uyz, p.z, acc[18]);
        acc[19] = fmaf(uzz, p.z, acc[19]);
        const float vx = w * x.x, vy = w * x.y, vz = w * x.z;
        acc[20] += vx; acc[21] += vy; acc[22] += vz;
        acc[23] = fmaf(vx, x.x, acc[23]); acc[24] = fmaf(vx, x.y, acc[24]); acc[25] = fmaf(vx, x.z, acc[25]);
        acc[26] = fmaf(vy, x.y, acc[26]); acc[27] = fmaf(vy, x.z, acc[27]); acc[28] = fmaf(vz, x.z, acc[28]);
        acc[29] = fmaf(vx, x.w, acc[29]); acc[30] = fmaf(vy, x.w, acc[30]); acc[31] = fmaf(vz, x.w, acc[31]);
      }
    }
    // Combine the G point lanes of every slot in fp64 (fixed order) into the CTA totals.
    constexpr int R = (F + 7) / 8;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (r > 0) __syncthreads();
      if (active) {
#pragma unroll
        for (int f = 0; f < 8; ++f)
          red[t * 8 + f] = (r * 8 + f < F) ? static_cast<double>(acc[(r * 8 + f < F) ? r * 8 + f : 0]) : 0.0;
      }
      __syncthreads();
      for (int o = t; o < K * 8; o += kTlsThreads) {
        const int kk = o >> 3, f = o & 7;
        if (r * 8 + f < F) {
          double s = 0.0;
          for (int gg = 0; gg < G; ++gg) s += red[(gg * K + kk) * 8 + f];
          tot[kk * kFP + r * 8 + f] += s;
        }
      }
    }
  }
  __syncthreads();
  double *out = part + (static_cast<size_t>(b) * chunks + chunk) * K * kFP;
  for (int o = t; o < K * kFP; o += kTlsThreads) {
    const int f = o & (kFP - 1);
    if (f >= F) continue;
    double v = tot[o];
    if (PASS == 1 && f >= 5 && f <= 13) {
      // features 5-13 were accumulated about the CTA's pivot c: back to the origin, in fp64 (exact algebra):
      //   sum w x = s + Sw c,   sum w x_i x_j = M_ij + c_i s_j + s_i c_j + Sw c_i c_j
      const double *row = tot + (o - f);
      const double sw = row[0];
      const double c[3] = {static_cast<double>(s_piv[0]), static_cast<double>(s_piv[1]), static_cast<double>(s_piv[2])};
      if (f < 8) {
        v += sw * c[f - 5];
      } else {
        const int ij[6][2] = {{0, 0}, {0, 1}, {0, 2}, {1, 1}, {1, 2}, {2, 2}};
        const int i = ij[f - 8][0], j = ij[f - 8][1];
        v += c[i] * row[5 + j] + row[5 + i] * c[j] + sw * c[i] * c[j];
      }
    }
    out[o] = v;
  }
}

// ---- small dense algebra in registers (fp64) -------------------------------------------------

// Cyclic Jacobi for a symmetric 3x3 (a = xx,xy,xz,yy,yz,zz).  lam[i], V[r][i] = i-th eigenpair.
__device__ void eig_sym3(const double a[6], double lam[3], double V[3][3]) {
  double A[3][3] = {{a[0], a[1], a[2]}, {a[1], a[3], a[4]}, {a[2], a[4], a[5]}};
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 12; ++sweep) {
    const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
    const double diag = fabs(A[0][0]) + fabs(A[1][1]) + fabs(A[2][2]);
    if (off <= 1e-18 * diag || off == 0.0) break;
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
      for (int q = p + 1; q < 3; ++q) {
        if (A[p][q] == 0.0) continue;
        const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
        const double tt = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(tt * tt + 1.0), s = tt * c;
#pragma unroll
        for (int r = 0; r < 3; ++r) {   // A <- A J
          const double arp = A[r][p], arq = A[r][q];
          A[r][p] = c * arp - s * arq; A[r][q] = s * arp + c * arq;
        }
#pragma unroll
        for (int r = 0; r < 3; ++r) {   // A <- J^T A
          const double apr = A[p][r], aqr = A[q][r];
          A[p][r] = c * apr - s * aqr; A[q][r] = s * apr + c * aqr;
        }
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const double vrp = V[r][p], vrq = V[r][q];
          V[r][p] = c * vrp - s * vrq; V[r][q] = s * vrp + c * vrq;
        }
      }
  }
  lam[0] = A[0][0]; lam[1] = A[1][1]; lam[2] = A[2][2];
}

// Right-singular vector of the smallest singular value of a symmetric matrix = eigenvector
// of the eigenvalue smallest in magnitude (ties -> last, like the last column of V).
__device__ void pick_min_eigvec(const double lam[3], const double V[3][3], double n[3]) {
  int m = 2;
  if (fabs(lam[1]) < fabs(lam[m])) m = 1;
  if (fabs(lam[0]) < fabs(lam[m])) m = 0;
  double x = m == 0 ? V[0][0] : (m == 1 ? V[0][1] : V[0][2]);
  double y = m == 0 ? V[1][0] : (m == 1 ? V[1][1] : V[1][2]);
  double z = m == 0 ? V[2][0] : (m == 1 ? V[2][1] : V[2][2]);
  const double nn = sqrt(x * x + y * y + z * z);
  if (nn > 0.0) { x /= nn; y /= nn; z /= nn; }
  // deterministic sign: the component of largest magnitude is positive
  const double ax = fabs(x), ay = fabs(y), az = fabs(z);
  const double lead = (ax >= ay && ax >= az) ? x : (ay >= az ? y : z);
  const double sg = lead < 0.0 ? -1.0 : 1.0;
  n[0] = sg * x; n[1] = sg * y; n[2] = sg * z;
}

// guarded_matrix_solve_ls (SPFN/geometry_utils.py:121-142) on the normal equations, given the
// eigenvalues of AtA: mask = cond(AtA) < 1e5 (singular values = |eigenvalues|), then
// (AtA*mask + 1e-8 I) x = Atb*mask by cofactors in fp64.
__device__ void guarded_solve3_eigs(const double a[6], const double rhs[3], const double lam[3], double x[3]) {
  const double s0 = fabs(lam[0]), s1 = fabs(lam[1]), s2 = fabs(lam[2]);
  const double smax = fmax(s0, fmax(s1, s2)), smin = fmin(s0, fmin(s1, s2));
  const double mask = (smax / smin < 1e5) ? 1.0 : 0.0;   // NaN / inf compare false
  const double m00 = a[0] * mask + 1e-8, m01 = a[1] * mask, m02 = a[2] * mask, m11 = a[3] * mask + 1e-8,
               m12 = a[4] * mask, m22 = a[5] * mask + 1e-8;
  const double r0 = rhs[0] * mask, r1 = rhs[1] * mask, r2 = rhs[2] * mask;
  const double c00 = m11 * m22 - m12 * m12, c01 = m02 * m12 - m01 * m22, c02 = m01 * m12 - m02 * m11;
  const double c11 = m00 * m22 - m02 * m02, c12 = m01 * m02 - m00 * m12, c22 = m00 * m11 - m01 * m01;
  const double det = m00 * c00 + m01 * c01 + m02 * c02;
  x[0] = (c00 * r0 + c01 * r1 + c02 * r2) / det;
  x[1] = (c01 * r0 + c11 * r1 + c12 * r2) / det;
  x[2] = (c02 * r0 + c12 * r1 + c22 * r2) / det;
}

__device__ void guarded_solve2(double axx, double axy, double ayy, const double rhs[2], double x[2]) {
  const double h = 0.5 * (axx + ayy), dlt = sqrt(0.25 * (axx - ayy) * (axx - ayy) + axy * axy);
  const double s0 = fabs(h + dlt), s1 = fabs(h - dlt);
  const double mask = (fmax(s0, s1) / fmin(s0, s1) < 1e5) ? 1.0 : 0.0;
  const double a = axx * mask + 1e-8, bb = axy * mask, d = ayy * mask + 1e-8;
  const double r0 = rhs[0] * mask, r1 = rhs[1] * mask;
  const double det = a * d - bb * bb;
  x[0] = (r0 * d - bb * r1) / det;
  x[1] = (a * r1 - bb * r0) / det;
}

__device__ __forceinline__ double sym6(const double a[6], int i, int j) {
  const int idx[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
  return a[idx[i][j]];
}

// index of T_ijk in (xxx,xxy,xxz,xyy,xyz,xzz,yyy,yyz,yzz,zzz)
__device__ __forceinline__ double sym10(const double t[10], int i, int j, int k) {
  int a = i, b = j, c = k, tmp;
  if (a > b) { tmp = a; a = b; b = tmp; }
  if (b > c) { tmp = b; b = c; c = tmp; }
  if (a > b) { tmp = a; a = b; b = tmp; }
  const int map[3][3][3] = {{{0, 1, 2}, {1, 3, 4}, {2, 4, 5}},
                            {{1, 3, 4}, {3, 6, 7}, {4, 7, 8}},
                            {{2, 4, 5}, {4, 7, 8}, {5, 8, 9}}};
  return t[map[a][b][c]];
}

// Sum the per-CTA partials of one (b,k): lane f owns feature f (coalesced 256-byte rows).
__device__ __forceinline__ double sum_partials(const double *part, int b, int k, int K, int chunks,
                                               int lane) {
  const double *p = part + (static_cast<size_t>(b) * chunks * K + k) * kFP + lane;
  // fixed summation order, eight independent loads in flight
  const size_t stride = static_cast<size_t>(K) * kFP;
  double s = 0.0;
  int c = 0;
  for (; c + 8 <= chunks; c += 8) {
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = p[(c + u) * stride];
#pragma unroll
    for (int u = 0; u < 8; ++u) s += v[u];
  }
  for (; c < chunks; ++c) s += p[c * stride];
  return s;
}

constexpr int kSolveWarps = 4;

__global__ void __launch_bounds__(kSolveWarps * 32)
tls_solve1_kernel(const double *__restrict__ part, double *__restrict__ state, int BK, int K, int chunks) {
  __shared__ double sm[kSolveWarps][kFP];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bk = blockIdx.x * kSolveWarps + warp;
  if (bk >= BK) return;
  const int b = bk / K, k = bk - b * K;
  sm[warp][lane] = lane < kF1 ? sum_partials(part, b, k, K, chunks, lane) : 0.0;
  __syncwarp();
  const double *m = sm[warp];
  double *st = state + static_cast<size_t>(bk) * kState;
  // The three 3x3 eigen problems of a slot run on three lanes of the SAME branch (in lock step), not one
  // after the other: lane 1 = cylinder axis (TLS on the normals, uncentred), lane 2 = cone apex, lane 3 = cone
  // axis (plane-fit normal of the normals: the smallest eigenvector of their scatter about the mean).
  double lam[3], V[3][3], Cx[6];
  {
    // sum w (x - mx)(x - mx)^T with the reference's fp32-rounded mean mx, from the raw moments (exact algebra)
    const double sw = m[0];
    const double denom = static_cast<double>(fmaxf(static_cast<float>(sw), 1e-10f));
    double mx[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) mx[i] = static_cast<double>(static_cast<float>(m[5 + i] / denom));
    const int ij[6][2] = {{0, 0}, {0, 1}, {0, 2}, {1, 1}, {1, 2}, {2, 2}};
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      const int i = ij[q][0], j = ij[q][1];
      Cx[q] = m[8 + q] - mx[i] * m[5 + j] - m[5 + i] * mx[j] + sw * mx[i] * mx[j];
    }
  }
  if (lane >= 1 && lane <= 3) eig_sym3(lane == 1 ? m + 8 : (lane == 2 ? m + 14 : Cx), lam, V);
  if (lane == 0) {
    const double sw = m[0];
    const double denom = static_cast<double>(fmaxf(static_cast<float>(sw), 1e-10f));
    st[ST_SW] = sw; st[ST_DENOM] = denom;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const double mu = static_cast<double>(static_cast<float>(m[1 + i] / denom));   // the reference's mean is fp32
      st[ST_MU + i] = mu;
      st[ST_MUX + i] = static_cast<double>(static_cast<float>(m[5 + i] / denom));
      st[ST_S1 + i] = m[1 + i] - mu * sw;                                             // sum w (p - mu)
    }
    st[ST_M2] = m[4] / denom;
  } else if (lane == 1) {
    double n[3];
    pick_min_eigvec(lam, V, n);
    st[ST_CYLN] = n[0]; st[ST_CYLN + 1] = n[1]; st[ST_CYLN + 2] = n[2];
  } else if (lane == 2) {
    double apex[3];
    guarded_solve3_eigs(m + 14, m + 20, lam, apex);   // rows sqrt(w') x, rhs sqrt(w') (p.x)
    st[ST_APEX] = apex[0]; st[ST_APEX + 1] = apex[1]; st[ST_APEX + 2] = apex[2];
  } else if (lane == 3) {
    double ax[3];
    pick_min_eigvec(lam, V, ax);                      // sign fixed in solve2 (cone_fitter.py:29-30)
#pragma unroll
    for (int i = 0; i < 3; ++i) st[ST_CONEAX + i] = static_cast<double>(static_cast<float>(ax[i]));
  }
}

__global__ void __launch_bounds__(kSolveWarps * 32)
tls_solve2_kernel(const double *__restrict__ part, double *__restrict__ state,
                  float *__restrict__ out, int BK, int K, int chunks) {
  __shared__ double sm[kSolveWarps][kFP];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bk = blockIdx.x * kSolveWarps + warp;
  if (bk >= BK) return;
  const int b = bk / K, k = bk - b * K;
  sm[warp][lane] = lane < kF2 ? sum_partials(part, b, k, K, chunks, lane) : 0.0;
  __syncwarp();
  if (lane > 3) return;
  const double *m = sm[warp];
  const double *S = m, *Sp = m + 6, *s1p = m + 12, *T = m + 15;
  double *st = state + static_cast<size_t>(bk) * kState;
  const double sw = st[ST_SW], denom = st[ST_DENOM], m2 = st[ST_M2];
  const double mu[3] = {st[ST_MU], st[ST_MU + 1], st[ST_MU + 2]};
  const double s1[3] = {st[ST_S1], st[ST_S1 + 1], st[ST_S1 + 2]};
  const size_t BKs = static_cast<size_t>(BK);
  float *o_pn = out, *o_pc = out + 3 * BKs, *o_sc = out + 4 * BKs, *o_sr = out + 7 * BKs,
        *o_ca = out + 8 * BKs, *o_cc = out + 11 * BKs, *o_cr = out + 14 * BKs,
        *o_ap = out + 15 * BKs, *o_ax = out + 18 * BKs;

  // The four sub-problems of a slot are independent: lanes 0..3 take one each.  The two 3x3 eigen
  // problems (plane S, sphere 4S') run in lock step on lanes 0 and 1 in ONE branch.
  double lam[3], V[3][3], AtA[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) AtA[i] = lane == 0 ? S[i] : 4.0 * Sp[i];
  if (lane < 2) eig_sym3(AtA, lam, V);
  if (lane == 0) {
    // plane: normal = TLS of the centred points, c = n . mean
    double n[3];
    pick_min_eigvec(lam, V, n);
#pragma unroll
    for (int i = 0; i < 3; ++i) o_pn[bk * 3 + i] = static_cast<float>(n[i]);
    o_pc[bk] = static_cast<float>(n[0] * mu[0] + n[1] * mu[1] + n[2] * mu[2]);
    return;
  }
  if (lane == 3) {
    // cone: apex and axis from solve1; sign so that the weighted mean of axis . dir is positive, 0 -> +1
    // (cone_fitter.py:29-30); half angle = sum w acos|axis . dir| / (sum w + 1e-10), clamped (:33-35)
    float *o_ha = out + 21 * BKs;
    const float sf = static_cast<float>(m[25]);
    const float sgn = sf > 0.f ? 1.f : (sf < 0.f ? -1.f : 1.f);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      o_ap[bk * 3 + i] = static_cast<float>(st[ST_APEX + i]);
      o_ax[bk * 3 + i] = static_cast<float>(st[ST_CONEAX + i]) * sgn;
    }
    float half = static_cast<float>(m[26]) / (static_cast<float>(sw) + 1e-10f);
    half = fminf(fmaxf(half, 1e-3f), 1.57079632679489661923f - 1e-3f);
    o_ha[bk] = half;
    return;
  }
  if (lane == 1) {
  // sphere: AtA = 4 S', Atb = -2 m2 s1' + 2 (|mu|^2 s1' + 2 S' mu + t'),  t'_i = sum_j T'_ijj
  const double mu2 = mu[0] * mu[0] + mu[1] * mu[1] + mu[2] * mu[2];
  double Atb[3], c[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double Smu = sym6(Sp, i, 0) * mu[0] + sym6(Sp, i, 1) * mu[1] + sym6(Sp, i, 2) * mu[2];
    const double t = sym10(T, i, 0, 0) + sym10(T, i, 1, 1) + sym10(T, i, 2, 2);
    Atb[i] = -2.0 * m2 * s1p[i] + 2.0 * (mu2 * s1p[i] + 2.0 * Smu + t);
  }
  guarded_solve3_eigs(AtA, Atb, lam, c);
  {
    const double e[3] = {mu[0] - c[0], mu[1] - c[1], mu[2] - c[2]};
    const double r2 = (S[0] + S[3] + S[5]) + 2.0 * (e[0] * s1[0] + e[1] * s1[1] + e[2] * s1[2]) +
                      (e[0] * e[0] + e[1] * e[1] + e[2] * e[2]) * sw;
#pragma unroll
    for (int i = 0; i < 3; ++i) o_sc[bk * 3 + i] = static_cast<float>(c[i]);
    o_sr[bk] = static_cast<float>(r2 / denom);
  }
  return;
  }

  // cylinder: frame of the axis (compute_consistent_plane_frame), circle fit from projected moments
  {
    const double a[3] = {st[ST_CYLN], st[ST_CYLN + 1], st[ST_CYLN + 2]};
    const double cand[3][3] = {{0.0, a[2], -a[1]}, {-a[2], 0.0, a[0]}, {a[1], -a[0], 0.0}};  // a x e_c
    int best = 0;
    double bn = -1.0;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const double nn = cand[q][0] * cand[q][0] + cand[q][1] * cand[q][1] + cand[q][2] * cand[q][2];
      if (nn > bn) { bn = nn; best = q; }
    }
    const double yn = fmax(sqrt(bn), 1e-12);
    const double ya[3] = {cand[best][0] / yn, cand[best][1] / yn, cand[best][2] / yn};
    const double xa[3] = {ya[1] * a[2] - ya[2] * a[1], ya[2] * a[0] - ya[0] * a[2], ya[0] * a[1] - ya[1] * a[0]};
    const double *Fm[2] = {xa, ya};
    double Sq[2][2], Spq[2][2], muq[2], s1q[2], s1pq[2], tq[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      muq[u] = Fm[u][0] * mu[0] + Fm[u][1] * mu[1] + Fm[u][2] * mu[2];
      s1q[u] = Fm[u][0] * s1[0] + Fm[u][1] * s1[1] + Fm[u][2] * s1[2];
      s1pq[u] = Fm[u][0] * s1p[0] + Fm[u][1] * s1p[1] + Fm[u][2] * s1p[2];
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        double acc = 0.0, accp = 0.0;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            acc += Fm[u][i] * sym6(S, i, j) * Fm[v][j];
            accp += Fm[u][i] * sym6(Sp, i, j) * Fm[v][j];
          }
        Sq[u][v] = acc; Spq[u][v] = accp;
      }
      double t = 0.0;
#pragma unroll
      for (int v = 0; v < 2; ++v)
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int l = 0; l < 3; ++l) t += Fm[u][i] * Fm[v][j] * Fm[v][l] * sym10(T, i, j, l);
      tq[u] = t;
    }
    const double muq2 = muq[0] * muq[0] + muq[1] * muq[1];
    const double m2q = ((Sq[0][0] + Sq[1][1]) + 2.0 * (muq[0] * s1q[0] + muq[1] * s1q[1]) + sw * muq2) / denom;
    double rhs[2], cq[2];
#pragma unroll
    for (int u = 0; u < 2; ++u)
      rhs[u] = -2.0 * m2q * s1pq[u] + 2.0 * (muq2 * s1pq[u] + 2.0 * (Spq[u][0] * muq[0] + Spq[u][1] * muq[1]) + tq[u]);
    guarded_solve2(4.0 * Spq[0][0], 4.0 * Spq[0][1], 4.0 * Spq[1][1], rhs, cq);
    const double e[2] = {muq[0] - cq[0], muq[1] - cq[1]};
    const double r2 = (Sq[0][0] + Sq[1][1]) + 2.0 * (e[0] * s1q[0] + e[1] * s1q[1]) + (e[0] * e[0] + e[1] * e[1]) * sw;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      o_ca[bk * 3 + i] = static_cast<float>(a[i]);
      o_cc[bk * 3 + i] = static_cast<float>(cq[0] * xa[i] + cq[1] * ya[i]);
    }
    o_cr[bk] = static_cast<float>(r2 / denom);
  }
}

// ---- training path: raw weighted moments and their gradients -----------------------------------
// M[b,k,f] = sum_n Wt[b,n,k] * psi_f(p_n, x_n), psi = [1, p(3), p p^T(6), p p p(10), x(3), x x^T(6), x (p.x)(3)].
// Linear in the weights, so the backward is dWt[b,n,k] = sum_f psi_f(n) dM[b,k,f] (same thread mapping as
// the forward: coalesced over k) and dX[b,n,:] = sum_k Wt[b,n,k] * sum_{f>=20} dM[b,k,f] dpsi_f/dx.
__global__ void __launch_bounds__(kSolveWarps * 32)
tls_sum_partials_kernel(const double *__restrict__ part, double *__restrict__ out, int BK, int K, int chunks) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bk = blockIdx.x * kSolveWarps + warp;
  if (bk >= BK) return;
  const int b = bk / K, k = bk - b * K;
  out[static_cast<size_t>(bk) * kFP + lane] = sum_partials(part, b, k, K, chunks, lane);
}

__global__ void __launch_bounds__(kTlsThreads)
moments_grad_w_kernel(const float *__restrict__ P, const float *__restrict__ X, const double *__restrict__ dM,
                      int N, int K, int G, int CP, float *__restrict__ dW) {
  extern __shared__ float4 s_pts[];
  float4 *sp = s_pts, *sx = s_pts + CP;
  const int b = blockIdx.y, t = threadIdx.x;
  const bool active = t < G * K;
  const int k = active ? t % K : 0, g = t / K;
  const int n0 = blockIdx.x * CP;
  const int cn = imin(CP, N - n0);
  float d[kF4];
  const double *dm = dM + (static_cast<size_t>(b) * K + k) * kFP;
#pragma unroll
  for (int f = 0; f < kF4; ++f) d[f] = static_cast<float>(dm[f]);
  for (int i = t; i < cn; i += kTlsThreads) {
    const float *p = P + (static_cast<size_t>(b) * N + n0 + i) * 3;
    const float *x = X + (static_cast<size_t>(b) * N + n0 + i) * 3;
    const float px = __ldg(p), py = __ldg(p + 1), pz = __ldg(p + 2);
    const float xx = __ldg(x), xy = __ldg(x + 1), xz = __ldg(x + 2);
    sp[i] = make_float4(px, py, pz, 0.f);
    sx[i] = make_float4(xx, xy, xz, px * xx + py * xy + pz * xz);
  }
  __syncthreads();
  if (!active) return;
  float *o = dW + (static_cast<size_t>(b) * N + n0) * K + k;
  for (int j = g; j < cn; j += G) {
    const float4 p = sp[j], x = sx[j];
    float a = d[0] + d[1] * p.x + d[2] * p.y + d[3] * p.z;
    const float pxx = p.x * p.x, pxy = p.x * p.y, pxz = p.x * p.z, pyy = p.y * p.y, pyz = p.y * p.z, pzz = p.z * p.z;
    a += d[4] * pxx + d[5] * pxy + d[6] * pxz + d[7] * pyy + d[8] * pyz + d[9] * pzz;
    a += d[10] * pxx * p.x + d[11] * pxx * p.y + d[12] * pxx * p.z + d[13] * pxy * p.y + d[14] * pxy * p.z +
         d[15] * pxz * p.z + d[16] * pyy * p.y + d[17] * pyy * p.z + d[18] * pyz * p.z + d[19] * pzz * p.z;
    a += d[20] * x.x + d[21] * x.y + d[22] * x.z;
    a += d[23] * x.x * x.x + d[24] * x.x * x.y + d[25] * x.x * x.z + d[26] * x.y * x.y + d[27] * x.y * x.z + d[28] * x.z * x.z;
    a += (d[29] * x.x + d[30] * x.y + d[31] * x.z) * x.w;
    o[static_cast<size_t>(j) * K] = a;
  }
}

// one thread per point: dX[n,:] = sum_k w[n,k] * (dMx + D x + dMxp (p.x) + (dMxp.x) p)
__global__ void __launch_bounds__(kTlsThreads)
moments_grad_x_kernel(const float *__restrict__ P, const float *__restrict__ X, const float *__restrict__ W,
                      const double *__restrict__ dM, int N, int K, float *__restrict__ dX) {
  extern __shared__ float s_dm[];            // [K][12]: features 20..31 of this cloud
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < K * 12; i += kTlsThreads)
    s_dm[i] = static_cast<float>(dM[(static_cast<size_t>(b) * K + i / 12) * kFP + 20 + i % 12]);
  __syncthreads();
  const int n = blockIdx.x * kTlsThreads + threadIdx.x;
  if (n >= N) return;
  const size_t row = static_cast<size_t>(b) * N + n;
  const float px = __ldg(P + row * 3), py = __ldg(P + row * 3 + 1), pz = __ldg(P + row * 3 + 2);
  const float xx = __ldg(X + row * 3), xy = __ldg(X + row * 3 + 1), xz = __ldg(X + row * 3 + 2);
  const float pdx = px * xx + py * xy + pz * xz;
  float gx = 0.f, gy = 0.f, gz = 0.f;
  const float *w = W + row * K;
  for (int k = 0; k < K; ++k) {
    const float *d = s_dm + k * 12;          // 0-2: x, 3-8: xx xy xz yy yz zz, 9-11: x (p.x)
    const float wk = __ldg(w + k);
    const float s = d[9] * xx + d[10] * xy + d[11] * xz;
    const float ax = d[0] + 2.f * d[3] * xx + d[4] * xy + d[5] * xz + d[9] * pdx + s * px;
    const float ay = d[1] + d[4] * xx + 2.f * d[6] * xy + d[7] * xz + d[10] * pdx + s * py;
    const float az = d[2] + d[5] * xx + d[7] * xy + 2.f * d[8] * xz + d[11] * pdx + s * pz;
    gx = fmaf(wk, ax, gx); gy = fmaf(wk, ay, gy); gz = fmaf(wk, az, gz);
  }
  dX[row * 3] = gx; dX[row * 3 + 1] = gy; dX[row * 3 + 2] = gz;
}

struct TlsWs {
  double *state, *part;
  size_t bytes;
};

TlsWs tls_carve(void *ws, int B, int K, const TlsGeom &g) {
  TlsWs w;
  const size_t n_state = static_cast<size_t>(B) * K * kState;
  const size_t n_part = static_cast<size_t>(B) * g.chunks * K * kFP;
  w.state = static_cast<double *>(ws);
  w.part = w.state + n_state;
  w.bytes = (n_state + n_part) * sizeof(double);
  return w;
}

}  // namespace
}  // namespace cpfn

extern "C" size_t cpfn_fit_workspace_bytes(int B, int N, int K) {
  using namespace cpfn;
  if (B <= 0 || N <= 0 || K <= 0 || K > kTlsThreads) return 0;
  const int sms = sm_count() > 0 ? sm_count() : 148;
  return tls_carve(nullptr, B, K, tls_geom(B, N, K, sms)).bytes;
}

extern "C" int cpfn_fit_primitives(const float *P, const float *W, const float *X, int B, int N,
                                   int K, float *out, void *workspace, size_t workspace_bytes,
                                   cpfn_stream_t stream) {
  using namespace cpfn;
  if (B < 0 || N < 0 || K < 0) return CPFN_EINVAL;
  if (B == 0 || K == 0) return CPFN_OK;
  if (N == 0 || K > kTlsThreads || !P || !W || !X || !out || B > 65535) return CPFN_EINVAL;
  const int sms = sm_count() > 0 ? sm_count() : 148;
  const TlsGeom g = tls_geom(B, N, K, sms);
  const TlsWs ws = tls_carve(workspace, B, K, g);
  if (!workspace || workspace_bytes < ws.bytes) return CPFN_EWORKSPACE;
  cudaStream_t st = as_stream(stream);
  const size_t smem = 2u * static_cast<size_t>(g.CP) * sizeof(float4) +
                      static_cast<size_t>(kTlsThreads) * 8 * sizeof(double) +
                      static_cast<size_t>(K) * kFP * sizeof(double);
  if (smem > 48 * 1024) {
    CPFN_CUDA_TRY(cudaFuncSetAttribute(tls_pass_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    CPFN_CUDA_TRY(cudaFuncSetAttribute(tls_pass_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  }
  const dim3 grid(g.chunks, B);
  const int BK = B * K;
  const int sgrid = (BK + kSolveWarps - 1) / kSolveWarps;
  tls_pass_kernel<1><<<grid, kTlsThreads, smem, st>>>(P, X, W, ws.state, ws.part, N, K, g.G, g.CP,
                                                      g.iters, g.chunks);
  tls_solve1_kernel<<<sgrid, kSolveWarps * 32, 0, st>>>(ws.part, ws.state, BK, K, g.chunks);
  tls_pass_kernel<2><<<grid, kTlsThreads, smem, st>>>(P, nullptr, W, ws.state, ws.part, N, K, g.G, g.CP,
                                                      g.iters, g.chunks);              // pass 2 reads no normals
  tls_solve2_kernel<<<sgrid, kSolveWarps * 32, 0, st>>>(ws.part, ws.state, out, BK, K, g.chunks);
  return check_launch();
}

extern "C" size_t cpfn_moments_workspace_bytes(int B, int N, int K) { return cpfn_fit_workspace_bytes(B, N, K); }

extern "C" int cpfn_weighted_moments(const float *P, const float *X, const float *Wt, int B, int N, int K,
                                     double *M, void *workspace, size_t workspace_bytes, cpfn_stream_t stream) {
  using namespace cpfn;
  if (B < 0 || N < 0 || K < 0) return CPFN_EINVAL;
  if (B == 0 || K == 0) return CPFN_OK;
  if (N == 0 || K > kTlsThreads || !P || !Wt || !X || !M || B > 65535) return CPFN_EINVAL;
  const int sms = sm_count() > 0 ? sm_count() : 148;
  const TlsGeom g = tls_geom(B, N, K, sms);
  const TlsWs ws = tls_carve(workspace, B, K, g);
  if (!workspace || workspace_bytes < ws.bytes) return CPFN_EWORKSPACE;
  cudaStream_t st = as_stream(stream);
  const size_t smem = 2u * static_cast<size_t>(g.CP) * sizeof(float4) +
                      static_cast<size_t>(kTlsThreads) * 8 * sizeof(double) + static_cast<size_t>(K) * kFP * sizeof(double);
  if (smem > 48 * 1024)
    CPFN_CUDA_TRY(cudaFuncSetAttribute(tls_pass_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  tls_pass_kernel<4><<<dim3(g.chunks, B), kTlsThreads, smem, st>>>(P, X, Wt, ws.state, ws.part, N, K, g.G, g.CP, g.iters,
                                                               g.chunks);
  const int BK = B * K;
  tls_sum_partials_kernel<<<(BK + kSolveWarps - 1) / kSolveWarps, kSolveWarps * 32, 0, st>>>(ws.part, M, BK, K, g.chunks);
  return check_launch();
}

extern "C" int cpfn_weighted_moments_grad(const float *P, const float *X, const float *Wt, const double *dM, int B,
                                          int N, int K, float *dWt, float *dX, cpfn_stream_t stream) {
  using namespace cpfn;
  if (B < 0 || N < 0 || K < 0) return CPFN_EINVAL;
  if (B == 0 || K == 0 || N == 0) return CPFN_OK;
  if (K > kTlsThreads || !P || !Wt || !X || !dM || B > 65535) return CPFN_EINVAL;
  cudaStream_t st = as_stream(stream);
  if (dWt) {
    const int G = imax(1, kTlsThreads / K);
    const int CP = G * imin(32, imax(1, kMaxCP / G));
    const size_t smem = 2u * static_cast<size_t>(CP) * sizeof(float4);
    if (smem > 48 * 1024)
      CPFN_CUDA_TRY(cudaFuncSetAttribute(moments_grad_w_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    moments_grad_w_kernel<<<dim3((N + CP - 1) / CP, B), kTlsThreads, smem, st>>>(P, X, dM, N, K, G, CP, dWt);
  }
  if (dX)
    moments_grad_x_kernel<<<dim3((N + kTlsThreads - 1) / kTlsThreads, B), kTlsThreads, static_cast<size_t>(K) * 12 * sizeof(float), st>>>(
        P, X, Wt, dM, N, K, dX);
  return check_launch();
}
