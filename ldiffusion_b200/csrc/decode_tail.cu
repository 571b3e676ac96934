// a-3 decode tail: planar fp32/bf16 image -> uint8 RGB (HWC) + PIL "L" gray (sm_100a).
//
// One pass over the decoder output: 3 planar streams in (128-bit loads), 16
// pixels per thread, one 16-byte gray store straight into the caller's slot of
// the [B, n+1, H, W] pixel-vector tensor and (optionally) three 16-byte stores
// of interleaved RGB.  HBM-bound: 3*s bytes in, 1 (+3) bytes out per pixel.
//
// Arithmetic (integer-exact against numpy/PIL):
//   t = clamp(x/2 + 0.5, 0, 1)            (in the image dtype: bf16 rounds per op,
//                                           as the reference's bf16 VAE would)
//   q = uint8(rint(fl32(t) * 255))         numpy round-half-even
//   L = (19595 R + 38470 G + 7471 B + 0x8000) >> 16
#include "common.cuh"

namespace ldiff {

template <typename T> __device__ __forceinline__ float round_storage(float x);
template <> __device__ __forceinline__ float round_storage<float>(float x) { return x; }
template <> __device__ __forceinline__ float round_storage<__nv_bfloat16>(float x) {
  return __bfloat162float(__float2bfloat16_rn(x));
}

template <typename T>
__device__ __forceinline__ uint32_t quantise(float x) {
  float t = round_storage<T>(__fmul_rn(x, 0.5f));
  t = round_storage<T>(__fadd_rn(t, 0.5f));
  t = fminf(fmaxf(t, 0.f), 1.f);
  return (uint32_t)__float2int_rn(__fmul_rn(t, 255.f));
}

__device__ __forceinline__ uint32_t luma(uint32_t r, uint32_t g, uint32_t b) {
  return (19595u * r + 38470u * g + 7471u * b + 0x8000u) >> 16;
}

template <typename T, bool RGB, bool GRAY>
__global__ void __launch_bounds__(256)
decode_tail_vec16_kernel(const T* __restrict__ img, uint8_t* __restrict__ rgb,
                         uint8_t* __restrict__ gray, int64_t hw, int64_t groups_per_img,
                         int64_t total_groups, int64_t gray_batch_stride) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t gidx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; gidx < total_groups;
       gidx += stride) {
    const int64_t b = gidx / groups_per_img;
    const int64_t p = (gidx - b * groups_per_img) << 4;
    const T* base = img + b * 3 * hw + p;
    uint32_t q[3][16];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float v[16];
      Vec8<T>::load(base + c * hw, reinterpret_cast<float(&)[8]>(v[0]));
      Vec8<T>::load(base + c * hw + 8, reinterpret_cast<float(&)[8]>(v[8]));
#pragma unroll
      for (int i = 0; i < 16; ++i) q[c][i] = quantise<T>(v[i]);
    }
    if (GRAY) {
      uint32_t w[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        w[j] = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          w[j] |= luma(q[0][4 * j + k], q[1][4 * j + k], q[2][4 * j + k]) << (8 * k);
      }
      __stcs(reinterpret_cast<uint4*>(gray + b * gray_batch_stride + p),
             make_uint4(w[0], w[1], w[2], w[3]));
    }
    if (RGB) {
      uint32_t w[12];
#pragma unroll
      for (int j = 0; j < 12; ++j) {
        w[j] = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int byte = 4 * j + k;
          w[j] |= q[byte % 3][byte / 3] << (8 * k);
        }
      }
      uint4* o = reinterpret_cast<uint4*>(rgb + (b * hw + p) * 3);
      __stcs(o, make_uint4(w[0], w[1], w[2], w[3]));
      __stcs(o + 1, make_uint4(w[4], w[5], w[6], w[7]));
      __stcs(o + 2, make_uint4(w[8], w[9], w[10], w[11]));
    }
  }
}

// any H*W (no alignment assumptions): one pixel per thread
template <typename T>
__global__ void __launch_bounds__(256)
decode_tail_scalar_kernel(const T* __restrict__ img, uint8_t* __restrict__ rgb,
                          uint8_t* __restrict__ gray, int64_t hw, int64_t total,
                          int64_t gray_batch_stride) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t b = i / hw, p = i - b * hw;
    const T* base = img + b * 3 * hw + p;
    const uint32_t r = quantise<T>(to_f32(base[0]));
    const uint32_t g = quantise<T>(to_f32(base[hw]));
    const uint32_t bl = quantise<T>(to_f32(base[2 * hw]));
    if (gray) gray[b * gray_batch_stride + p] = (uint8_t)luma(r, g, bl);
    if (rgb) {
      uint8_t* o = rgb + i * 3;
      o[0] = (uint8_t)r; o[1] = (uint8_t)g; o[2] = (uint8_t)bl;
    }
  }
}

template <typename T>
static int launch_decode_tail(const void* img, uint8_t* rgb, uint8_t* gray, int B, int H, int W,
                              int64_t gray_batch_stride, cudaStream_t st) {
  const int64_t hw = (int64_t)H * W;
  const int threads = 256;
  const bool vec_ok = (hw % 16 == 0) && aligned16(img) && aligned16(rgb) && aligned16(gray) &&
                      (gray_batch_stride % 16 == 0);
  if (vec_ok) {
    const int64_t gpi = hw >> 4, total = gpi * B;
    const int grid = grid_for(total, threads, 8);
    const T* p = (const T*)img;
    if (rgb && gray)
      decode_tail_vec16_kernel<T, true, true><<<grid, threads, 0, st>>>(p, rgb, gray, hw, gpi, total, gray_batch_stride);
    else if (gray)
      decode_tail_vec16_kernel<T, false, true><<<grid, threads, 0, st>>>(p, rgb, gray, hw, gpi, total, gray_batch_stride);
    else
      decode_tail_vec16_kernel<T, true, false><<<grid, threads, 0, st>>>(p, rgb, gray, hw, gpi, total, gray_batch_stride);
  } else {
    const int64_t total = hw * B;
    decode_tail_scalar_kernel<T><<<grid_for(total, threads, 8), threads, 0, st>>>(
        (const T*)img, rgb, gray, hw, total, gray_batch_stride);
  }
  return check_launch();
}

}  // namespace ldiff

using namespace ldiff;

extern "C" int ldiff_decode_tail_gray(const void* img, uint8_t* rgb_hwc, uint8_t* gray, int B, int H,
                                      int W, int64_t gray_batch_stride, int dtype, void* stream) {
  if (!img || (!rgb_hwc && !gray) || B < 0 || H < 0 || W < 0) return LDIFF_EINVAL;
  if (gray && gray_batch_stride < (int64_t)H * W) return LDIFF_EINVAL;
  if (B == 0 || H == 0 || W == 0) return LDIFF_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == LDIFF_F32) return launch_decode_tail<float>(img, rgb_hwc, gray, B, H, W, gray_batch_stride, st);
  if (dtype == LDIFF_BF16)
    return launch_decode_tail<__nv_bfloat16>(img, rgb_hwc, gray, B, H, W, gray_batch_stride, st);
  return LDIFF_EUNSUPPORTED;
}
