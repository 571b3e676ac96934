// Widening N2: nnU-Net's sliding-window accumulation, mirror-TTA merge and
// softmax/argmax export on the tissue path (sm_100a).
//
// Reference (vendored nnU-Net v2.6.2): predict_from_raw_data.py:530-545 (mirror TTA),
// :547-589 (gaussian-weighted accumulation in torch.half + normalisation),
// label_handling.py:128-173 (logits.float() -> softmax(0) -> argmax(0)).
// The results arrays are torch.half there, so every eager op rounds to fp16; the kernels
// reproduce that chain of roundings (compute in fp32, round to half after each reference op),
// which makes them bit-identical to the eager chain.
#include <cuda_fp16.h>

#include "common.cuh"

namespace ldiff {

__device__ __forceinline__ float h2f(__half h) { return __half2float(h); }
__device__ __forceinline__ __half f2h(float f) { return __float2half_rn(f); }

// predicted_logits[sl] += prediction * gaussian ; n_predictions[sl[1:]] += gaussian
// (predict_from_raw_data.py:577-578).  One thread per tile pixel, loop over heads.
__global__ void __launch_bounds__(256)
sw_accumulate_kernel(const __half* __restrict__ pred, const __half* __restrict__ gauss,
                     __half* __restrict__ acc, __half* __restrict__ npred, int K, int th, int tw,
                     int H, int W, int y0, int x0) {
  const int n = th * tw;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int ty = i / tw, tx = i - ty * tw;
    const int64_t o = (int64_t)(y0 + ty) * W + (x0 + tx);
    const float g = gauss ? h2f(gauss[i]) : 1.f;
    for (int k = 0; k < K; ++k) {
      const float p = h2f(pred[(int64_t)k * n + i]);
      const __half term = gauss ? f2h(__fmul_rn(p, g)) : f2h(p);          // prediction * gaussian (half)
      __half* a = acc + (int64_t)k * H * W + o;
      *a = f2h(__fadd_rn(h2f(*a), h2f(term)));                            // += (half)
    }
    npred[o] = f2h(__fadd_rn(h2f(npred[o]), g));
  }
}

// prediction += flip(network(flip(x))) for every axes combination, then /= (n + 1)
// (predict_from_raw_data.py:541-544).  preds[0] is the unflipped prediction; preds[j] is the
// network output on the input flipped along flips[j] (bit 0: rows, bit 1: columns).
struct TtaList { const __half* p[8]; int flip[8]; int n; };

__global__ void __launch_bounds__(256)
sw_tta_merge_kernel(TtaList list, __half* __restrict__ out, int K, int th, int tw) {
  const int n = th * tw;
  const int total = K * n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int k = i / n, r = i - k * n;
    const int y = r / tw, x = r - y * tw;
    __half s = list.p[0][i];
    for (int j = 1; j < list.n; ++j) {
      const int yy = (list.flip[j] & 1) ? th - 1 - y : y;
      const int xx = (list.flip[j] & 2) ? tw - 1 - x : x;
      s = f2h(__fadd_rn(h2f(s), h2f(list.p[j][(int64_t)k * n + (int64_t)yy * tw + xx])));
    }
    out[i] = f2h(__fdiv_rn(h2f(s), (float)list.n));
  }
}

// predicted_logits /= n_predictions (half), then logits.float() -> softmax(0) -> argmax(0).
// Decision rule as in head.cu: first argmax unless another head is within 1e-5, then the pinned
// softmax (exp in fp64 -> fp32, sequential fp32 sum, IEEE division, first maximum).
constexpr int kMaxHeads = 32;

__global__ void __launch_bounds__(256)
sw_finalize_argmax_kernel(const __half* __restrict__ acc, const __half* __restrict__ npred,
                          uint8_t* __restrict__ seg, __half* __restrict__ logits_out, int K, int64_t hw,
                          int* __restrict__ status) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += (int64_t)gridDim.x * blockDim.x) {
    const float nn = h2f(npred[i]);
    float v[kMaxHeads];
    float best = -INFINITY, second = -INFINITY;
    int idx = 0;
    bool bad = false;
    for (int k = 0; k < K; ++k) {
      const __half q = f2h(__fdiv_rn(h2f(acc[(int64_t)k * hw + i]), nn));
      if (logits_out) logits_out[(int64_t)k * hw + i] = q;
      const float x = h2f(q);
      v[k] = x;
      bad |= isinf(x);
      if (x > best) { second = best; best = x; idx = k; }
      else second = fmaxf(second, x);
    }
    if (bad) atomicOr(status, LDIFF_STATUS_SW_INF);      // the reference raises on inf (:581-585)
    if (K > 1 && __fsub_rn(best, second) <= 1e-5f) {
      float s = 0.f;
      for (int k = 0; k < K; ++k) s = __fadd_rn(s, (float)exp((double)__fsub_rn(v[k], best)));
      float pb = -1.f;
      for (int k = 0; k < K; ++k) {
        const float p = __fdiv_rn((float)exp((double)__fsub_rn(v[k], best)), s);
        if (p > pb) { pb = p; idx = k; }
      }
    }
    seg[i] = (uint8_t)idx;
  }
}

// ---------------------------------------------------------------------------------------
// 8-wide (16-byte) variants: same arithmetic, eight fp16 values per load/store
// ---------------------------------------------------------------------------------------
struct H8 { __half2 h[4]; };
static_assert(sizeof(H8) == 16, "H8 is one 128-bit vector");

__device__ __forceinline__ H8 ld8(const __half* p) {
  const uint4 r = *reinterpret_cast<const uint4*>(p);
  return *reinterpret_cast<const H8*>(&r);
}
__device__ __forceinline__ void st8(__half* p, const H8& v) {
  *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(&v);
}
__device__ __forceinline__ void unpack8(const H8& v, float (&f)[8]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 t = __half22float2(v.h[j]);
    f[2 * j] = t.x; f[2 * j + 1] = t.y;
  }
}
__device__ __forceinline__ H8 pack8(const float (&f)[8]) {
  H8 v;
#pragma unroll
  for (int j = 0; j < 4; ++j) v.h[j] = __floats2half2_rn(f[2 * j], f[2 * j + 1]);
  return v;
}

__global__ void __launch_bounds__(256)
sw_accumulate_vec8_kernel(const __half* __restrict__ pred, const __half* __restrict__ gauss,
                          __half* __restrict__ acc, __half* __restrict__ npred, int K, int th, int tw,
                          int H, int W, int y0, int x0) {
  const int gw = tw >> 3, ngroups = th * gw, n = th * tw;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ngroups; i += gridDim.x * blockDim.x) {
    const int ty = i / gw, tx = (i - ty * gw) << 3;
    const int t = ty * tw + tx;
    const int64_t o = (int64_t)(y0 + ty) * W + (x0 + tx);
    float g[8];
    if (gauss) unpack8(ld8(gauss + t), g);
    else {
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] = 1.f;
    }
    for (int k0 = 0; k0 < K; k0 += 4) {                // four heads per step: eight 128-bit loads in flight
      H8 pv[4], av[4];
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (k0 + q < K) {
          pv[q] = ld8(pred + (int64_t)(k0 + q) * n + t);
          av[q] = ld8(acc + (int64_t)(k0 + q) * H * W + o);
        }
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (k0 + q < K) {
          float p[8], a[8];
          unpack8(pv[q], p);
          unpack8(av[q], a);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float term = gauss ? h2f(f2h(__fmul_rn(p[j], g[j]))) : p[j];
            a[j] = __fadd_rn(a[j], term);
          }
          st8(acc + (int64_t)(k0 + q) * H * W + o, pack8(a));
        }
    }
    float nv[8];
    unpack8(ld8(npred + o), nv);
#pragma unroll
    for (int j = 0; j < 8; ++j) nv[j] = __fadd_rn(nv[j], g[j]);
    st8(npred + o, pack8(nv));
  }
}

__global__ void __launch_bounds__(256)
sw_tta_merge_vec8_kernel(TtaList list, __half* __restrict__ out, int K, int th, int tw) {
  const int gw = tw >> 3, n = th * tw;
  const int total = K * th * gw;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int k = i / (th * gw), r = i - k * th * gw;
    const int y = r / gw, x = (r - y * gw) << 3;
    float s[8];
    unpack8(ld8(list.p[0] + (int64_t)k * n + (int64_t)y * tw + x), s);
    for (int j = 1; j < list.n; ++j) {
      const int yy = (list.flip[j] & 1) ? th - 1 - y : y;
      const bool fx = (list.flip[j] & 2) != 0;
      const int xx = fx ? tw - 8 - x : x;              // mirrored vector, elements reversed below
      float o[8];
      unpack8(ld8(list.p[j] + (int64_t)k * n + (int64_t)yy * tw + xx), o);
#pragma unroll
      for (int e = 0; e < 8; ++e) s[e] = h2f(f2h(__fadd_rn(s[e], fx ? o[7 - e] : o[e])));
    }
    const float inv_n = (float)list.n;
#pragma unroll
    for (int e = 0; e < 8; ++e) s[e] = __fdiv_rn(s[e], inv_n);
    st8(out + (int64_t)k * n + (int64_t)y * tw + x, pack8(s));
  }
}

// eight pixels per thread; the running (best, second, index) triples live in registers
__global__ void __launch_bounds__(256)
sw_finalize_argmax_vec8_kernel(const __half* __restrict__ acc, const __half* __restrict__ npred,
                               uint8_t* __restrict__ seg, __half* __restrict__ logits_out, int K, int64_t hw,
                               int* __restrict__ status) {
  const int64_t ngroups = hw >> 3;
  for (int64_t gi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; gi < ngroups;
       gi += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = gi << 3;
    float nn[8], best[8], second[8];
    int idx[8];
    unpack8(ld8(npred + i), nn);
#pragma unroll
    for (int e = 0; e < 8; ++e) { best[e] = -INFINITY; second[e] = -INFINITY; idx[e] = 0; }
    bool bad = false;
    for (int k0 = 0; k0 < K; k0 += 4) {                // four heads per step: four 128-bit loads in flight
      H8 av[4];
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (k0 + q < K) av[q] = ld8(acc + (int64_t)(k0 + q) * hw + i);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (k0 + q < K) {
          const int k = k0 + q;
          float a[8];
          unpack8(av[q], a);
#pragma unroll
          for (int e = 0; e < 8; ++e) a[e] = __fdiv_rn(a[e], nn[e]);
          const H8 qv = pack8(a);
          if (logits_out) st8(logits_out + (int64_t)k * hw + i, qv);
          unpack8(qv, a);                              // the half-rounded logits, as floats
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            bad |= isinf(a[e]);
            if (a[e] > best[e]) { second[e] = best[e]; best[e] = a[e]; idx[e] = k; }
            else second[e] = fmaxf(second[e], a[e]);
          }
        }
    }
    if (bad) atomicOr(status, LDIFF_STATUS_SW_INF);
    uint32_t lo = 0, hi = 0, ties = 0;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (K > 1 && __fsub_rn(best[e], second[e]) <= 1e-5f) ties |= 1u << e;
      if (e < 4) lo |= (uint32_t)idx[e] << (8 * e);
      else hi |= (uint32_t)idx[e] << (8 * (e - 4));
    }
    while (ties) {                                     // tie / near-tie: pinned softmax (re-reads the K values)
      const int e = __ffs(ties) - 1;
      ties &= ties - 1;
      const float nne = h2f(npred[i + e]);
      float m = -INFINITY;
      for (int k = 0; k < K; ++k) m = fmaxf(m, h2f(f2h(__fdiv_rn(h2f(acc[(int64_t)k * hw + i + e]), nne))));
      float s = 0.f;
      for (int k = 0; k < K; ++k) {
        const float x = h2f(f2h(__fdiv_rn(h2f(acc[(int64_t)k * hw + i + e]), nne)));
        s = __fadd_rn(s, (float)exp((double)__fsub_rn(x, m)));
      }
      float pb = -1.f;
      int r = 0;
      for (int k = 0; k < K; ++k) {
        const float x = h2f(f2h(__fdiv_rn(h2f(acc[(int64_t)k * hw + i + e]), nne)));
        const float p = __fdiv_rn((float)exp((double)__fsub_rn(x, m)), s);
        if (p > pb) { pb = p; r = k; }
      }
      if (e < 4) lo = (lo & ~(0xffu << (8 * e))) | ((uint32_t)r << (8 * e));
      else hi = (hi & ~(0xffu << (8 * (e - 4)))) | ((uint32_t)r << (8 * (e - 4)));
    }
    *reinterpret_cast<uint2*>(seg + i) = make_uint2(lo, hi);
  }
}

}  // namespace ldiff

using namespace ldiff;

extern "C" int ldiff_sw_accumulate(const void* pred_f16, const void* gauss_f16, void* acc_f16, void* npred_f16,
                                   int K, int th, int tw, int H, int W, int y0, int x0, void* stream) {
  if (!pred_f16 || !acc_f16 || !npred_f16 || K < 1 || th < 1 || tw < 1 || H < 1 || W < 1) return LDIFF_EINVAL;
  if (y0 < 0 || x0 < 0 || y0 + th > H || x0 + tw > W) return LDIFF_EINVAL;
  const bool vec = (tw % 8 == 0) && (W % 8 == 0) && (x0 % 8 == 0) && aligned16(pred_f16) && aligned16(gauss_f16) &&
                   aligned16(acc_f16) && aligned16(npred_f16);
  if (vec)
    sw_accumulate_vec8_kernel<<<grid_for((int64_t)th * tw / 8, 256, 8), 256, 0, (cudaStream_t)stream>>>(
        (const __half*)pred_f16, (const __half*)gauss_f16, (__half*)acc_f16, (__half*)npred_f16, K, th, tw, H, W,
        y0, x0);
  else
    sw_accumulate_kernel<<<grid_for((int64_t)th * tw, 256, 8), 256, 0, (cudaStream_t)stream>>>(
        (const __half*)pred_f16, (const __half*)gauss_f16, (__half*)acc_f16, (__half*)npred_f16, K, th, tw, H, W,
        y0, x0);
  return check_launch();
}

extern "C" int ldiff_sw_tta_merge(const void* const* host_preds_f16, const int* host_flips, int n, void* out_f16,
                                  int K, int th, int tw, void* stream) {
  if (!host_preds_f16 || !host_flips || n < 1 || n > 8 || !out_f16 || K < 1 || th < 1 || tw < 1) return LDIFF_EINVAL;
  TtaList list{};
  list.n = n;
  for (int j = 0; j < n; ++j) {
    if (!host_preds_f16[j] || (host_flips[j] & ~3)) return LDIFF_EINVAL;
    list.p[j] = (const __half*)host_preds_f16[j];
    list.flip[j] = host_flips[j];
  }
  if (list.flip[0] != 0) return LDIFF_EINVAL;
  bool vec = (tw % 8 == 0) && aligned16(out_f16);
  for (int j = 0; j < n; ++j) vec = vec && aligned16(host_preds_f16[j]);
  if (vec)
    sw_tta_merge_vec8_kernel<<<grid_for((int64_t)K * th * tw / 8, 256, 8), 256, 0, (cudaStream_t)stream>>>(
        list, (__half*)out_f16, K, th, tw);
  else
    sw_tta_merge_kernel<<<grid_for((int64_t)K * th * tw, 256, 8), 256, 0, (cudaStream_t)stream>>>(
        list, (__half*)out_f16, K, th, tw);
  return check_launch();
}

extern "C" int ldiff_sw_finalize_argmax(const void* acc_f16, const void* npred_f16, uint8_t* seg, void* logits_out_f16,
                                        int K, int64_t hw, int* status, void* stream) {
  if (!acc_f16 || !npred_f16 || !seg || !status || K < 1 || hw < 0) return LDIFF_EINVAL;
  if (K > kMaxHeads) return LDIFF_EUNSUPPORTED;
  if (hw == 0) return LDIFF_OK;
  const bool vec = (hw % 8 == 0) && aligned16(acc_f16) && aligned16(npred_f16) && aligned16(logits_out_f16) &&
                   ((reinterpret_cast<uintptr_t>(seg) & 7u) == 0);
  if (vec)
    sw_finalize_argmax_vec8_kernel<<<grid_for(hw / 8, 256, 8), 256, 0, (cudaStream_t)stream>>>(
        (const __half*)acc_f16, (const __half*)npred_f16, seg, (__half*)logits_out_f16, K, hw, status);
  else
    sw_finalize_argmax_kernel<<<grid_for(hw, 256, 8), 256, 0, (cudaStream_t)stream>>>(
        (const __half*)acc_f16, (const __half*)npred_f16, seg, (__half*)logits_out_f16, K, hw, status);
  return check_launch();
}
