// a-6 confusion matrix C[(K+1),K] from uint8 prediction / ground-truth maps (sm_100a).
//
// One pass, 2 bytes per pixel: each thread streams 16 pixels per iteration with
// two 128-bit loads, merges equal neighbours into runs in registers (label maps
// are spatially coherent, so most 4-pixel words extend the current run with two
// compares), and flushes a run with ONE shared-memory atomic into a histogram
// that is replicated per lane (hist[bin][lane], bank == lane: a warp-wide flush
// never bank-conflicts and never contends inside the warp).  At the end the
// block folds the replicas and issues one 64-bit global atomic per non-empty
// bin into the caller's int64 matrix (which may be the NCCL all-reduce buffer).
#include "common.cuh"

namespace ldiff {

template <int R>
__global__ void __launch_bounds__(512)
confusion_hist_kernel(const uint8_t* __restrict__ pred, const uint8_t* __restrict__ gt,
                      const uint8_t* __restrict__ gt_lut, unsigned long long* __restrict__ C,
                      int64_t n, int K, int* __restrict__ status) {
  extern __shared__ uint32_t hist[];                 // [nbins][R]
  __shared__ uint8_t lut[256];
  const int nbins = (K + 1) * K;
  for (int i = threadIdx.x; i < nbins * R; i += blockDim.x) hist[i] = 0;
  for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = gt_lut ? gt_lut[i] : (uint8_t)i;
  __syncthreads();

  const int rep = threadIdx.x & (R - 1);
  int bad = 0;
  uint32_t run_p = 0, run_g = 0, run_cnt = 0;

  auto flush = [&]() {
    if (run_cnt) {
      const int row = min((int)lut[run_g], K);
      if ((int)run_p < K) atomicAdd(&hist[(row * K + (int)run_p) * R + rep], run_cnt);
      else bad = 1;
    }
  };
  auto word = [&](uint32_t pw, uint32_t gw) {
    if (pw == run_p * 0x01010101u && gw == run_g * 0x01010101u) {
      run_cnt += 4;
      return;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t p = (pw >> (8 * k)) & 0xffu, g = (gw >> (8 * k)) & 0xffu;
      if (p == run_p && g == run_g) {
        ++run_cnt;
      } else {
        flush();
        run_p = p; run_g = g; run_cnt = 1;
      }
    }
  };

  const int64_t nvec = n >> 4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const uint4* pv = reinterpret_cast<const uint4*>(pred);
  const uint4* gv = reinterpret_cast<const uint4*>(gt);
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
    const uint4 a = __ldcs(pv + v);
    const uint4 b = __ldcs(gv + v);
    word(a.x, b.x); word(a.y, b.y); word(a.z, b.z); word(a.w, b.w);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {          // n % 16 trailing pixels
    for (int64_t i = nvec << 4; i < n; ++i) {
      const uint32_t p = pred[i], g = gt[i];
      if (p == run_p && g == run_g) ++run_cnt;
      else { flush(); run_p = p; run_g = g; run_cnt = 1; }
    }
  }
  flush();
  if (bad) atomicOr(status, LDIFF_STATUS_PRED_RANGE);
  __syncthreads();

  for (int bin = threadIdx.x; bin < nbins; bin += blockDim.x) {
    unsigned long long s = 0;
#pragma unroll 8
    for (int r = 0; r < R; ++r) s += hist[bin * R + ((r + threadIdx.x) & (R - 1))];
    if (s) atomicAdd(C + bin, s);
  }
}

__global__ void __launch_bounds__(256)
labels_to_u8_kernel(const int64_t* __restrict__ in, uint8_t* __restrict__ out, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t nvec = n >> 2;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
    const longlong2 a = __ldcs(reinterpret_cast<const longlong2*>(in) + 2 * v);
    const longlong2 b = __ldcs(reinterpret_cast<const longlong2*>(in) + 2 * v + 1);
    const long long x[4] = {a.x, a.y, b.x, b.y};
    uint32_t w = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) w |= (uint32_t)((x[k] < 0 || x[k] > 254) ? 255 : x[k]) << (8 * k);
    reinterpret_cast<uint32_t*>(out)[v] = w;
  }
  const int64_t t = (nvec << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) out[t] = (uint8_t)((in[t] < 0 || in[t] > 254) ? 255 : in[t]);
}

}  // namespace ldiff

using namespace ldiff;

extern "C" int ldiff_confusion_hist(const uint8_t* pred, const uint8_t* gt, const uint8_t* gt_lut,
                                    int64_t* C, int64_t n, int K, int* status, void* stream) {
  if (!pred || !gt || !C || !status || n < 0 || K < 1) return LDIFF_EINVAL;
  if (K > 128) return LDIFF_EUNSUPPORTED;
  if (n == 0) return LDIFF_OK;
  if (!aligned16(pred) || !aligned16(gt)) return LDIFF_EALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  const int threads = 512;
  const int nbins = (K + 1) * K;
  int R = 32;
  while (R > 1 && (size_t)nbins * R * 4 > 64 * 1024) R >>= 1;
  const size_t smem = (size_t)nbins * R * 4;
  const int grid = grid_for((n >> 4) > 0 ? (n >> 4) : 1, threads, 4);
  unsigned long long* Cu = reinterpret_cast<unsigned long long*>(C);
#define CH(RR)                                                                                      \
  do {                                                                                              \
    static bool attr_set = false;                                                                   \
    if (!attr_set && smem > 48 * 1024) {                                                            \
      cudaFuncSetAttribute(confusion_hist_kernel<RR>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                           72 * 1024);                                                              \
      attr_set = true;                                                                              \
    }                                                                                               \
    confusion_hist_kernel<RR><<<grid, threads, smem, st>>>(pred, gt, gt_lut, Cu, n, K, status);     \
  } while (0)
  switch (R) {
    case 32: CH(32); break;
    case 16: CH(16); break;
    case 8: CH(8); break;
    case 4: CH(4); break;
    case 2: CH(2); break;
    default: CH(1); break;
  }
#undef CH
  return check_launch();
}

extern "C" int ldiff_labels_to_u8(const int64_t* in, uint8_t* out, int64_t n, void* stream) {
  if (!in || !out || n < 0) return LDIFF_EINVAL;
  if (n == 0) return LDIFF_OK;
  if (!aligned16(in) || (reinterpret_cast<uintptr_t>(out) & 3u)) return LDIFF_EALIGN;
  const int64_t items = (n >> 2) > (n & 3) ? (n >> 2) : (n & 3);
  labels_to_u8_kernel<<<grid_for(items, 256, 8), 256, 0, (cudaStream_t)stream>>>(in, out, n);
  return check_launch();
}
