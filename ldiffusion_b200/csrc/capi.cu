// Library-level entry points of libldiff_sm100.so.
#include <cstdlib>

#include "common.cuh"

namespace ldiff {

unsigned long long g_launches = 0;

int current_device() {
  int dev = -1;
  return cudaGetDevice(&dev) == cudaSuccess ? dev : -1;
}

// SM count of the CURRENT device; one cached slot per device ordinal (a process may drive several GPUs)
int sm_count() {
  static int cached[kMaxDevices] = {};
  const int dev = current_device();
  if (dev < 0 || dev >= kMaxDevices) return 148;      // B200
  int n = __atomic_load_n(&cached[dev], __ATOMIC_RELAXED);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    __atomic_store_n(&cached[dev], n, __ATOMIC_RELAXED);
  }
  return n;
}

// tuning knobs: -1 = not set yet (the environment variable, else the built-in default, is taken on first use)
static int g_tune[LDIFF_TUNE_COUNT] = {-1, -1, -1, -1};
static const char* const kTuneEnv[LDIFF_TUNE_COUNT] = {"LDIFF_ARGMAX_VARIANT", "LDIFF_DT_SMS", "LDIFF_DT_TMA",
                                                        "LDIFF_PHILOX_ROUNDS"};
static const int kTuneDefault[LDIFF_TUNE_COUNT] = {0, 0, 6, 0};

int tune_get(int knob) {
  int v = __atomic_load_n(&g_tune[knob], __ATOMIC_RELAXED);
  if (v < 0) {
    const char* e = getenv(kTuneEnv[knob]);
    v = e ? atoi(e) : kTuneDefault[knob];
    if (v < 0) v = 0;
    __atomic_store_n(&g_tune[knob], v, __ATOMIC_RELAXED);
  }
  return v;
}

// dst plane b (dst + b*dst_stride) = src plane b (src + b*n), n bytes each, 16 bytes per thread
__global__ void __launch_bounds__(256)
copy_planes_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int64_t n, int64_t dst_stride) {
  const uint4* s = reinterpret_cast<const uint4*>(src + (int64_t)blockIdx.y * n);
  uint4* d = reinterpret_cast<uint4*>(dst + (int64_t)blockIdx.y * dst_stride);
  const int64_t nvec = n >> 4;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (int64_t)gridDim.x * blockDim.x)
    __stcs(d + v, __ldcs(s + v));
}

}  // namespace ldiff

// label plane of the pixel vectors (pixel_latent_vector.py:92): B planes of n bytes into a strided slot
extern "C" int ldiff_copy_planes_u8(const uint8_t* src, uint8_t* dst, int64_t n, int B, int64_t dst_stride,
                                    void* stream) {
  if (!src || !dst || n < 0 || B < 0 || dst_stride < n) return LDIFF_EINVAL;
  if (n == 0 || B == 0) return LDIFF_OK;
  if (!ldiff::aligned16(src) || !ldiff::aligned16(dst) || (n % 16) || (dst_stride % 16)) return LDIFF_EALIGN;
  if (B > 65535) return LDIFF_EUNSUPPORTED;
  int64_t bx = ((n >> 4) + 255) / 256;
  const int64_t cap = ((int64_t)ldiff::sm_count() * 8 + B - 1) / B;
  if (bx > cap) bx = cap;
  ldiff::copy_planes_kernel<<<dim3((unsigned)bx, (unsigned)B), 256, 0, (cudaStream_t)stream>>>(src, dst, n,
                                                                                             dst_stride);
  return ldiff::check_launch();
}

extern "C" int ldiff_tune(int knob, int value) {
  if (knob < 0 || knob >= LDIFF_TUNE_COUNT || value < 0) return LDIFF_EINVAL;
  __atomic_store_n(&ldiff::g_tune[knob], value, __ATOMIC_RELAXED);
  return LDIFF_OK;
}

extern "C" int ldiff_tune_get(int knob) {
  if (knob < 0 || knob >= LDIFF_TUNE_COUNT) return LDIFF_EINVAL;
  return ldiff::tune_get(knob);
}

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
