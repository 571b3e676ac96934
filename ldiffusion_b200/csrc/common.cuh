// Shared helpers for the sm_100a kernels of libldiff_sm100.so.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "ldiff.h"

namespace ldiff {

// ---- launch bookkeeping ----------------------------------------------------
extern unsigned long long g_launches;          // defined in capi.cu

constexpr int kMaxDevices = 64;
int current_device();                          // cudaGetDevice, -1 on error
int sm_count();                                // SM count of the CURRENT device (cached per device)

// ldiff_tune knobs (capi.cu); first use reads the environment (LDIFF_ARGMAX_VARIANT, LDIFF_DT_SMS, LDIFF_DT_TMA, ...)
int tune_get(int knob);

inline int check_launch() {
  __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED);
  return cudaPeekAtLastError() == cudaSuccess ? LDIFF_OK : LDIFF_ELAUNCH;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// grid for a grid-stride loop over `items` work items with `threads` per block:
// enough blocks to cover the work, capped at `per_sm` resident blocks per SM.
inline int grid_for(int64_t items, int threads, int per_sm) {
  int64_t need = (items + threads - 1) / threads;
  int64_t cap = (int64_t)sm_count() * per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

// ---- storage types -----------------------------------------------------------
// 8 consecutive elements of a floating tensor <-> 8 fp32 registers.
template <typename T> struct Vec8;

template <> struct Vec8<float> {
  static constexpr int kBytes = 32;
  __device__ static void load(const float* p, float (&v)[8]) {
    float4 a = __ldcs(reinterpret_cast<const float4*>(p));
    float4 b = __ldcs(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  __device__ static void store(float* p, const float (&v)[8]) {
    __stcs(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
    __stcs(reinterpret_cast<float4*>(p) + 1, make_float4(v[4], v[5], v[6], v[7]));
  }
};

template <> struct Vec8<__nv_bfloat16> {
  static constexpr int kBytes = 16;
  __device__ static void load(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 r = __ldcs(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  __device__ static void store(__nv_bfloat16* p, const float (&v)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    __stcs(reinterpret_cast<uint4*>(p), make_uint4(w[0], w[1], w[2], w[3]));
  }
};

// 16 consecutive instance ids of a label image as four int4 (the unit the LUT paint kernels work on), from int32
// ids or from the uint16 ids Cellpose hands over below 65 536 labels (conductor.py:180)
__device__ __forceinline__ void load_ids16(const int32_t* p, int4 (&q)[4]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) q[j] = __ldcs(reinterpret_cast<const int4*>(p) + j);
}
__device__ __forceinline__ void load_ids16(const uint16_t* p, int4 (&q)[4]) {
  const uint4 a = __ldcs(reinterpret_cast<const uint4*>(p)), b = __ldcs(reinterpret_cast<const uint4*>(p) + 1);
  const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int j = 0; j < 4; ++j)
    q[j] = make_int4((int)(w[2 * j] & 0xffffu), (int)(w[2 * j] >> 16), (int)(w[2 * j + 1] & 0xffffu), (int)(w[2 * j + 1] >> 16));
}

// Optional side job of a chain's FIRST kernel: zero the int64 counters (a confusion matrix) that a later
// kernel of the same stream accumulates into, so a pass needs no memset node that every chain waits for.
__device__ __forceinline__ void clear_counters(unsigned long long* __restrict__ p, int n) {
  if (p != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0)
    for (int i = threadIdx.x; i < n; i += blockDim.x) p[i] = 0ull;
}

__device__ __forceinline__ float to_f32(float x) { return x; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 x) { return __bfloat162float(x); }
__device__ __forceinline__ float to_f32(uint8_t x) { return (float)x; }

template <typename T> __device__ __forceinline__ T from_f32(float x);
template <> __device__ __forceinline__ float from_f32<float>(float x) { return x; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float x) {
  return __float2bfloat16_rn(x);
}
// float -> uint8 as torch's .to(torch.uint8) does for in-range values: truncate
template <> __device__ __forceinline__ uint8_t from_f32<uint8_t>(float x) {
  return (uint8_t)(int)x;
}

}  // namespace ldiff
