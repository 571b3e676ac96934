// Library-level entry points of libldiff_sm100.so.
#include "common.cuh"

namespace ldiff {

unsigned long long g_launches = 0;

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      cached = 148;                                   // B200
  }
  return cached;
}

}  // namespace ldiff

extern "C" int ldiff_abi_version(void) { return LDIFF_ABI_VERSION; }

extern "C" int64_t ldiff_launch_count(void) {
  return (int64_t)__atomic_load_n(&ldiff::g_launches, __ATOMIC_RELAXED);
}

extern "C" const char* ldiff_strerror(int code) {
  switch (code) {
    case LDIFF_OK: return "ok";
    case LDIFF_EINVAL: return "invalid argument";
    case LDIFF_EALIGN: return "pointer or stride not 16-byte aligned";
    case LDIFF_ELAUNCH: return "CUDA kernel launch failed";
    case LDIFF_EUNSUPPORTED: return "unsupported dtype / shape combination";
    default: return "unknown ldiff error";
  }
}
