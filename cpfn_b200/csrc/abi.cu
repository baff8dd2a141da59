// Library identity, error reporting and device queries of libcpfn_b200.
#include "common.cuh"

namespace cpfn {
namespace {
thread_local cudaError_t g_last = cudaSuccess;
}

void set_last_cuda_error(cudaError_t e) { g_last = e; }

int sm_count() {
  static thread_local int cached_dev = -1;
  static thread_local int cached = -1;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (dev != cached_dev) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    cached = v;
    cached_dev = dev;
  }
  return cached;
}
}  // namespace cpfn

extern "C" int cpfn_version(void) { return 100; }

extern "C" const char *cpfn_error_string(int code) {
  switch (code) {
    case CPFN_OK: return "ok";
    case CPFN_EINVAL: return "invalid argument (size, null pointer or unsupported shape)";
    case CPFN_ELAUNCH: return "CUDA launch failed (see cpfn_last_cuda_error)";
    case CPFN_EWORKSPACE: return "workspace missing or too small";
    default: return "unknown cpfn error";
  }
}

extern "C" const char *cpfn_last_cuda_error(void) { return cudaGetErrorString(cpfn::g_last); }

extern "C" int cpfn_sm_count(void) { return cpfn::sm_count(); }
