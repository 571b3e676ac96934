// Helpers shared by the a-5 kernels (head.cu, lift_argmax_env.cu): the bilinear taps with the reference's
// roundings and the pinned softmax/argmax decision rule (oracle/head.py::softmax_argmax_spec).
#pragma once
#include <utility>

#include "common.cuh"

namespace ldiff {

struct AxisH { float scale; int in, out; };
struct TapH { int i0, i1; float l0, l1; };

__device__ __forceinline__ TapH tap(const AxisH& a, int dst) {
  TapH t;
  if (a.in == a.out) { t.i0 = t.i1 = dst; t.l0 = 1.f; t.l1 = 0.f; return t; }
  float src = fmaxf(__fmaf_rn(a.scale, (float)dst + 0.5f, -0.5f), 0.f);
  t.i0 = min((int)floorf(src), a.in - 1);
  t.i1 = min(t.i0 + 1, a.in - 1);
  t.l1 = fminf(fmaxf(__fsub_rn(src, (float)t.i0), 0.f), 1.f);
  t.l0 = __fsub_rn(1.f, t.l1);
  return t;
}

__device__ __forceinline__ float lerp2(float w0, float a, float w1, float b) {
  return __fmaf_rn(w0, a, __fmul_rn(w1, b));
}

constexpr float kTieGap = 1e-5f;

// pinned softmax + first-max argmax over v[lo..K)
template <int KT>
static __device__ __noinline__ int softmax_argmax_exact(const float (&v)[KT], int K, int lo) {
  float m = v[0];
  for (int k = 1; k < K; ++k) m = fmaxf(m, v[k]);
  float e[KT];
  float s = 0.f;
  for (int k = 0; k < K; ++k) {
    e[k] = (float)exp((double)__fsub_rn(v[k], m));
    s = __fadd_rn(s, e[k]);
  }
  int best = lo;
  float pb = __fdiv_rn(e[lo], s);
  for (int k = lo + 1; k < K; ++k) {
    const float p = __fdiv_rn(e[k], s);
    if (p > pb) { pb = p; best = k; }
  }
  return best;
}

// cell form (conductor.py:218-221): softmax(...)[:, 1:] -> top-1 -> +1 is the first argmax over k >= 1 unless two
// logits are within kTieGap, in which case the pinned softmax decides.  v[k >= K] must be -inf.
template <int KT>
__device__ __forceinline__ int cell_decide(const float (&v)[KT], int K) {
  if (K <= 1) return 0;
  float best = v[1], second = -INFINITY;
  int cls = 1;
#pragma unroll
  for (int k = 2; k < KT; ++k)
    if (k < K) {
      if (v[k] > best) { second = best; best = v[k]; cls = k; }
      else second = fmaxf(second, v[k]);
    }
  if (__fsub_rn(best, second) <= kTieGap) cls = softmax_argmax_exact<KT>(v, K, 1);
  return cls;
}

// Cold path taken INSIDE the row loop: the K lifted logits of one pixel by value (the
// struct travels through the ABI's parameter space, so the hot loop keeps its
// register allocation), no memory traffic, ~1k instructions.
template <int K>
struct Vals { float v[K]; };

template <int K>
static __device__ __noinline__ int exact_from_values(Vals<K> x) {
  float m = x.v[0];
#pragma unroll 1
  for (int k = 1; k < K; ++k) m = fmaxf(m, x.v[k]);
  float e[K];
  float s = 0.f;
#pragma unroll 1
  for (int k = 0; k < K; ++k) {
    e[k] = (float)exp((double)__fsub_rn(x.v[k], m));
    s = __fadd_rn(s, e[k]);
  }
  int best = 0;
  float pb = __fdiv_rn(e[0], s);
#pragma unroll 1
  for (int k = 1; k < K; ++k) {
    const float p = __fdiv_rn(e[k], s);
    if (p > pb) { pb = p; best = k; }
  }
  return best;
}

// Cold path of the generic kernel: one pixel resolved from scratch with the pinned softmax.  Kept out of
// line and fed scalars only so that it costs the hot loops no registers.
static __device__ __noinline__ int exact_pixel(const float* __restrict__ lb, int K, int plane, int in_w,
                                        int yi0, int yi1, float yl0, float yl1, int xi0, int xi1,
                                        float xl0, float xl1) {
  auto value = [&](int k) {
    const float* r0 = lb + k * plane + yi0 * in_w;
    const float* r1 = lb + k * plane + yi1 * in_w;
    return lerp2(yl0, lerp2(xl0, __ldg(r0 + xi0), xl1, __ldg(r0 + xi1)), yl1,
                 lerp2(xl0, __ldg(r1 + xi0), xl1, __ldg(r1 + xi1)));
  };
  float m = value(0);
  for (int k = 1; k < K; ++k) m = fmaxf(m, value(k));
  float s = 0.f;
  for (int k = 0; k < K; ++k) s = __fadd_rn(s, (float)exp((double)__fsub_rn(value(k), m)));
  float pb = -1.f;
  int idx = 0;
  for (int k = 0; k < K; ++k) {
    const float pk = __fdiv_rn((float)exp((double)__fsub_rn(value(k), m)), s);
    if (pk > pb) { pb = pk; idx = k; }
  }
  return idx;
}


constexpr int kBand = 128;                             // longest band (output rows per source row)
constexpr int kQueue = 128;

// longest band of a lift h -> H: the first one, every output row whose upper tap is source row 0
// (including the rows clamped at src < 0): about 1.5 * H / h rows
inline int max_band_rows(int h, int H) { return (int)((3ll * H + 2ll * h - 1) / (2ll * h)) + 2; }

}  // namespace ldiff
