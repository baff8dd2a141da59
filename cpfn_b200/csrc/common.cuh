// Shared device/host helpers for libcpfn_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/cpfn_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libcpfn_b200 is written for sm_100a (B200) only"
#endif

namespace cpfn {

constexpr int kWarp = 32;

// Records the last CUDA failure of this host thread for cpfn_last_cuda_error().
void set_last_cuda_error(cudaError_t e);

inline int check_launch() {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_cuda_error(e);
    return CPFN_ELAUNCH;
  }
  return CPFN_OK;
}

#define CPFN_CUDA_TRY(expr)                 \
  do {                                      \
    cudaError_t _e = (expr);                \
    if (_e != cudaSuccess) {                \
      ::cpfn::set_last_cuda_error(_e);      \
      return CPFN_ELAUNCH;                  \
    }                                       \
  } while (0)

inline cudaStream_t as_stream(cpfn_stream_t s) {
  return reinterpret_cast<cudaStream_t>(s);
}

int sm_count();

// Squared distance with the reference's rounding sequence, read off the SASS of
// the reference kernels built for sm_100a: nvcc contracts x*x + y*y + z*z as
// FMUL(y*y), FFMA(x,x,.), FFMA(z,z,.).  The _rn intrinsics pin the sequence:
// nvcc neither contracts nor re-associates them.
__device__ __forceinline__ float sqdist3(float ax, float ay, float az, float bx,
                                         float by, float bz) {
  const float dx = __fsub_rn(ax, bx);
  const float dy = __fsub_rn(ay, by);
  const float dz = __fsub_rn(az, bz);
  return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

__device__ __forceinline__ float sqnorm3(float x, float y, float z) {
  return __fmaf_rn(z, z, __fmaf_rn(x, x, __fmul_rn(y, y)));
}

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

}  // namespace cpfn
