// a-1 Laplace forward noising and a-2 PLMS reverse step (sm_100a).
//
// Both are streaming elementwise kernels over a latent tensor: 128-bit loads
// and stores with streaming cache hints, 8 elements per thread per iteration,
// grid-stride over a grid sized to the SM count.  All arithmetic uses the _rn
// intrinsics in the order of the reference's eager op chain so that no FMA
// contraction changes a rounding (bit-exact against oracle/ in fp32).
#include "common.cuh"

namespace ldiff {

// ---------------------------------------------------------------------------
// Philox4x32-R (Salmon et al., SC'11), counter-based: no state in memory.  R = 10 is the published default
// (cuRAND's and torch's choice), R = 7 the smallest round count that passes BigCrush in the paper's table: 30 %
// fewer multiply rounds for the bf16 instantiation, whose kernel is bound by exactly those IMAD.WIDEs.
// ---------------------------------------------------------------------------
template <int R>
__device__ __forceinline__ uint4 philox4x32(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

// One 32-bit word -> one Laplace(0, b) variate.  Bits [22:0] are the magnitude m (23 bits), bit 31 the sign (each a
// single LOP3 on the word; bits [30:23] are not used):
//   |u| = m * 2^-23 in [0, 1)      (f = 1 + |u| is built from the bit pattern, no int->float conversion;
//   1 - |u| = 2 - f                 both subtractions are exact)
//   noise = sign * b * (-ln(1 - |u|))           == torch's  -b * sign(u) * log1p(-|u|)  with u = sign * |u|
// A symmetric 24-bit uniform: |u| never reaches 1 (the largest magnitude is b * 23 ln 2 = 15.9 b), u = 0 has
// probability 2^-23.  log: the bare MUFU.LG2 (1 - |u| >= 2^-23 is never denormal, so __log2f's FSETP/FMUL/FADD
// bracket is dead weight).
//  SERIES (fp32 storage): |u| >= 2^-5 uses the hardware log2 (relative error <= 6e-6 there), smaller |u| a
//   5-term series (relative error < 1e-8): the SAME variate as the libm chain to ~6e-6 relative, far inside the
//   1e-3 contract, at a third of libm's instruction count (libm made the fused kernel issue-bound at 0.59).
//  !SERIES (bf16 storage, where half the bytes move per element and the kernel is ALU-bound): the hardware
//   log2 everywhere (absolute error <= 1.7e-7 b, below a bf16 ulp of any noise value above 5e-5 b), sign and
//   scale folded into one signed multiplier.
__device__ __forceinline__ float lg2_normal(float w) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(w));
  return r;
}
template <bool SERIES>
__device__ __forceinline__ float laplace_from_word(uint32_t w, float b) {
  const float f = __uint_as_float(0x3f800000u | (w & 0x007fffffu));
  const float t = __fsub_rn(2.f, f);                                   // 1 - |u|, exact
  const uint32_t sgn = w & 0x80000000u;
  if (!SERIES)
    return lg2_normal(t) * __uint_as_float(__float_as_uint(b * -0.6931471805599453f) ^ sgn);
  const float a = __fadd_rn(f, -1.f);                                  // |u|, exact
  const float lg = lg2_normal(t) * 0.6931471805599453f;
  float p = __fmaf_rn(a, 0.2f, 0.25f);
  p = __fmaf_rn(a, p, 0.3333333333f);
  p = __fmaf_rn(a, p, 0.5f);
  p = __fmaf_rn(a, p, 1.f);
  const float l = (a < 0.03125f) ? -a * p : lg;                        // log1p(-a) <= 0
  return __uint_as_float(__float_as_uint(b * -l) ^ sgn);
}
template <typename T> struct FastLaplace { static constexpr bool kSeries = true; static constexpr bool kHalfWords = false; };
template <> struct FastLaplace<__nv_bfloat16> { static constexpr bool kSeries = false; static constexpr bool kHalfWords = true; };

// bf16 storage draws TWO variates from every 32-bit word (the kernel is bound by Philox's IMAD.WIDEs, and a bf16 sum
// cannot resolve a 23-bit uniform anyway): each half-word is a sign bit + a 15-bit magnitude m, |u| = m * 2^-15.
// The exponential tail is NOT truncated at 15 ln 2: by the memorylessness of -ln(1 - U), the top cell m = 2^15 - 1
// (probability 2^-15) is refined with 23 more bits of a second draw, |noise| = b * (15 ln 2 - ln(1 - m2 * 2^-23)), so
// the support reaches 38 ln 2 = 26.3 b (the 24-bit fp32 form stops at 15.9 b).
//   half 0 = bits [31:16] (sign 31, magnitude [30:16]), half 1 = bits [15:0] (sign 15, magnitude [14:0])
// lg2_of_half: log2(1 - |u|) of one half and its sign bit; the rare top cell is flagged for the caller.
__device__ __forceinline__ float lg2_of_half(uint32_t w, int half, uint32_t& sgn, bool& top) {
  const uint32_t m8 = (half == 0 ? (w >> 8) : (w << 8)) & 0x007fff00u;    // the 15 magnitude bits at mantissa [22:8]
  sgn = (half == 0 ? w : (w << 16)) & 0x80000000u;
  top = m8 == 0x007fff00u;
  return lg2_normal(__fsub_rn(2.f, __uint_as_float(0x3f800000u | m8)));   // 1 - |u| >= 2^-15: exact, normal
}
// second-level draw for element e (0..7) of the 8-element group whose counter is c: log2 of the refined tail
template <int R>
__device__ __noinline__ float lg2_tail_extension(uint64_t c, int e, uint2 key) {
  const uint4 r = philox4x32<R>(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 1u + (uint32_t)e, 0u), key);
  return -15.f + lg2_normal(__fsub_rn(2.f, __uint_as_float(0x3f800000u | (r.x & 0x007fffffu))));
}

// torch.distributions.Laplace.rsample with loc = 0:
//   noise = 0 - (scale * sign(u)) * log1p(-|u|)
__device__ __forceinline__ float laplace_from_uniform(float u, float b) {
  const float sgn = (u > 0.f) ? 1.f : ((u < 0.f) ? -1.f : 0.f);
  const float t = __fmul_rn(b, sgn);
  const float l = log1pf(-fabsf(u));
  return __fsub_rn(0.f, __fmul_rn(t, l));
}

enum { SRC_PHILOX = 0, SRC_UNIFORM = 1, SRC_NOISE = 2, SRC_PHILOX7 = 3 };   // SRC_PHILOX: 10 rounds
template <int SRC> struct IsPhilox { static constexpr bool value = SRC == SRC_PHILOX || SRC == SRC_PHILOX7; };
template <int SRC> struct Rounds { static constexpr int value = SRC == SRC_PHILOX7 ? 7 : 10; };

// rounds of the Philox stream: 10 (the published default) unless LDIFF_TUNE_PHILOX_ROUNDS = 7 asks for the
// paper's Crush-resistant minimum
template <typename T> static bool philox_seven() { return tune_get(LDIFF_TUNE_PHILOX_ROUNDS) == 7; }

// 8 elements of the half-word form: lg[i] = log2(1 - |u_i|) <= 0 (tail extension applied), sg[i] = sign bit
template <int R>
__device__ __forceinline__ void philox_half8(int64_t base, uint2 key, uint64_t offset, float (&lg)[8], uint32_t (&sg)[8]) {
  const uint64_t c = offset + (uint64_t)(base >> 3);
  const uint4 r = philox4x32<R>(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 0u, 0u), key);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
  bool any_top = false;
  bool top[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    lg[i] = lg2_of_half(w[i >> 1], i & 1, sg[i], top[i]);
    any_top |= top[i];
  }
  if (any_top) {                                       // probability 2^-12 per group
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (top[i]) lg[i] = lg2_tail_extension<R>(c, i, key);
  }
}

// the 8 Philox words of elements [base, base + 8)
template <int SRC>
__device__ __forceinline__ void philox_words8(int64_t base, uint2 key, uint64_t offset, uint32_t (&w)[8]) {
  constexpr int R = SRC == SRC_PHILOX7 ? 7 : 10;
  const uint64_t c0 = offset + (uint64_t)(base >> 2);
  const uint4 r0 = philox4x32<R>(make_uint4((uint32_t)c0, (uint32_t)(c0 >> 32), 0u, 0u), key);
  const uint64_t c1 = c0 + 1;
  const uint4 r1 = philox4x32<R>(make_uint4((uint32_t)c1, (uint32_t)(c1 >> 32), 0u, 0u), key);
  w[0] = r0.x; w[1] = r0.y; w[2] = r0.z; w[3] = r0.w; w[4] = r1.x; w[5] = r1.y; w[6] = r1.z; w[7] = r1.w;
}

// x + noise for the bf16-storage Philox form when the noise itself is not emitted: the scale-and-sign multiply and
// the add are ONE fma (the sum is rounded to bf16 right after, so the fp32 rounding of the noise it skips is 2^-16
// of a storage ulp; the emitting form keeps noise and sum separate so that out == x + noise_out exactly)
template <typename T, int SRC> struct FusedAdd {
  static constexpr bool value = IsPhilox<SRC>::value && FastLaplace<T>::kHalfWords;
};

// noise of the 8 elements [base, base + 8) / of element t (shared by the stand-alone kernel and the
// fused step+noise kernel, so both produce the same bits)
template <typename T, int SRC>
__device__ __forceinline__ void laplace_noise8(const T* __restrict__ inj, int64_t base, float b, uint2 key,
                                               uint64_t offset, float (&nz)[8]) {
  if (IsPhilox<SRC>::value && FastLaplace<T>::kHalfWords) {
    float lg[8];
    uint32_t sg[8];
    philox_half8<Rounds<SRC>::value>(base, key, offset, lg, sg);
    const uint32_t kb = __float_as_uint(b * -0.6931471805599453f);
#pragma unroll
    for (int i = 0; i < 8; ++i) nz[i] = lg[i] * __uint_as_float(kb ^ sg[i]);
  } else if (IsPhilox<SRC>::value) {
    uint32_t w[8];
    philox_words8<SRC>(base, key, offset, w);
#pragma unroll
    for (int i = 0; i < 8; ++i) nz[i] = laplace_from_word<FastLaplace<T>::kSeries>(w[i], b);
  } else {
    Vec8<T>::load(inj + base, nz);
    if (SRC == SRC_UNIFORM) {
#pragma unroll
      for (int i = 0; i < 8; ++i) nz[i] = laplace_from_uniform(nz[i], b);
    }
  }
}
template <typename T, int SRC>
__device__ __forceinline__ float laplace_noise1(const T* __restrict__ inj, int64_t t, float b, uint2 key,
                                                uint64_t offset) {
  if (IsPhilox<SRC>::value && FastLaplace<T>::kHalfWords) {
    constexpr int R = Rounds<SRC>::value;
    const uint64_t c = offset + (uint64_t)(t >> 3);
    const uint4 r = philox4x32<R>(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 0u, 0u), key);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
    const int e = (int)(t & 7);
    uint32_t sg;
    bool top;
    float lg = lg2_of_half(w[e >> 1], e & 1, sg, top);
    if (top) lg = lg2_tail_extension<R>(c, e, key);
    return lg * __uint_as_float(__float_as_uint(b * -0.6931471805599453f) ^ sg);
  }
  if (IsPhilox<SRC>::value) {
    constexpr int R = Rounds<SRC>::value;
    const uint64_t c = offset + (uint64_t)(t >> 2);
    const uint4 r = philox4x32<R>(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 0u, 0u), key);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
    return laplace_from_word<FastLaplace<T>::kSeries>(w[t & 3], b);
  }
  const float nzs = to_f32(inj[t]);
  return SRC == SRC_UNIFORM ? laplace_from_uniform(nzs, b) : nzs;
}

template <typename T, int SRC, bool EMIT>
__global__ void __launch_bounds__(256)
laplace_qsample_kernel(const T* __restrict__ x, T* __restrict__ out, const T* __restrict__ inj,
                       T* __restrict__ noise_out, float b, uint2 key, uint64_t offset, int64_t n) {
  const int64_t nvec = n >> 3;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  auto body = [&](int64_t base) {
    float xv[8], nz[8];
    Vec8<T>::load(x + base, xv);
    if (!EMIT && FusedAdd<T, SRC>::value) {
      float lg[8];
      uint32_t sg[8];
      philox_half8<Rounds<SRC>::value>(base, key, offset, lg, sg);
      const uint32_t kb = __float_as_uint(b * -0.6931471805599453f);
#pragma unroll
      for (int i = 0; i < 8; ++i) xv[i] = __fmaf_rn(lg[i], __uint_as_float(kb ^ sg[i]), xv[i]);
    } else {
      laplace_noise8<T, SRC>(inj, base, b, key, offset, nz);
      if (EMIT) Vec8<T>::store(noise_out + base, nz);
#pragma unroll
      for (int i = 0; i < 8; ++i) xv[i] = __fadd_rn(xv[i], nz[i]);
    }
    Vec8<T>::store(out + base, xv);
  };
  if (n < (int64_t(1) << 31)) {                          // 32-bit index arithmetic (the bf16 form is ALU-bound)
    const uint32_t nv = (uint32_t)nvec, st = (uint32_t)stride;
    for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += st) body((int64_t)(v << 3));
  } else {
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) body(v << 3);
  }
  // scalar tail (n % 8 elements), one thread each
  const int64_t t = (nvec << 3) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) {
    const float nzs = laplace_noise1<T, SRC>(inj, t, b, key, offset);
    if (EMIT) noise_out[t] = from_f32<T>(nzs);
    out[t] = from_f32<T>(__fadd_rn(to_f32(x[t]), nzs));
  }
}

template <typename T>
static int launch_qsample(const void* x, void* out, const void* noise_in, const void* u_in,
                          void* noise_out, float b, uint64_t seed, uint64_t offset, int64_t n,
                          cudaStream_t st) {
  const int threads = 256;
  const int64_t items = (n >> 3) > (n & 7) ? (n >> 3) : (n & 7);
  const int grid = grid_for(items, threads, 8);
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  const T* xs = (const T*)x;
  T* os = (T*)out;
  T* no = (T*)noise_out;
#define LQ(SRC, EMIT, INJ) \
  laplace_qsample_kernel<T, SRC, EMIT><<<grid, threads, 0, st>>>(xs, os, (const T*)(INJ), no, b, key, offset, n)
  if (noise_in) {
    if (no) LQ(SRC_NOISE, true, noise_in); else LQ(SRC_NOISE, false, noise_in);
  } else if (u_in) {
    if (no) LQ(SRC_UNIFORM, true, u_in); else LQ(SRC_UNIFORM, false, u_in);
  } else {
    if (philox_seven<T>()) { if (no) LQ(SRC_PHILOX7, true, nullptr); else LQ(SRC_PHILOX7, false, nullptr); }
    else if (no) LQ(SRC_PHILOX, true, nullptr); else LQ(SRC_PHILOX, false, nullptr);
  }
#undef LQ
  return check_launch();
}

// ---------------------------------------------------------------------------
// PLMS step: prev = sc * x - (dA * eh) / denom, eh = Adams-Bashforth combination
// ---------------------------------------------------------------------------
template <int MODE>
__device__ __forceinline__ float plms_eps(float e0, float e1, float e2, float e3) {
  if (MODE == 0) return e0;
  if (MODE == 1) return __fmul_rn(__fadd_rn(e0, e1), 0.5f);                       // (a + b) / 2
  if (MODE == 2) return __fmul_rn(__fsub_rn(__fmul_rn(3.f, e0), e1), 0.5f);       // (3a - b) / 2
  if (MODE == 3)
    return __fdiv_rn(__fadd_rn(__fsub_rn(__fmul_rn(23.f, e0), __fmul_rn(16.f, e1)),
                               __fmul_rn(5.f, e2)), 12.f);
  const float c24 = (float)(1.0 / 24.0);
  return __fmul_rn(c24, __fsub_rn(__fadd_rn(__fsub_rn(__fmul_rn(55.f, e0), __fmul_rn(59.f, e1)),
                                            __fmul_rn(37.f, e2)), __fmul_rn(9.f, e3)));
}

template <typename T, int MODE>
__global__ void __launch_bounds__(256)
plms_step_kernel(const T* __restrict__ x, const T* __restrict__ e0, const T* __restrict__ e1,
                 const T* __restrict__ e2, const T* __restrict__ e3, float sc, float dA,
                 float denom, T* __restrict__ out, int64_t n) {
  const int64_t nvec = n >> 3;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
    const int64_t base = v << 3;
    float xv[8], a[8], b[8], c[8], d[8];
    Vec8<T>::load(x + base, xv);
    Vec8<T>::load(e0 + base, a);
    if (MODE >= 1) Vec8<T>::load(e1 + base, b);
    if (MODE >= 3) Vec8<T>::load(e2 + base, c);
    if (MODE >= 4) Vec8<T>::load(e3 + base, d);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float eh = plms_eps<MODE>(a[i], MODE >= 1 ? b[i] : 0.f, MODE >= 3 ? c[i] : 0.f,
                                      MODE >= 4 ? d[i] : 0.f);
      xv[i] = __fsub_rn(__fmul_rn(sc, xv[i]), __fdiv_rn(__fmul_rn(dA, eh), denom));
    }
    Vec8<T>::store(out + base, xv);
  }
  const int64_t t = (nvec << 3) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) {
    const float eh = plms_eps<MODE>(to_f32(e0[t]), MODE >= 1 ? to_f32(e1[t]) : 0.f,
                                    MODE >= 3 ? to_f32(e2[t]) : 0.f, MODE >= 4 ? to_f32(e3[t]) : 0.f);
    out[t] = from_f32<T>(__fsub_rn(__fmul_rn(sc, to_f32(x[t])), __fdiv_rn(__fmul_rn(dA, eh), denom)));
  }
}

template <typename T>
static int launch_plms(const void* x, const void* e0, const void* e1, const void* e2, const void* e3,
                       int mode, float sc, float dA, float denom, void* out, int64_t n,
                       cudaStream_t st) {
  const int threads = 256;
  const int64_t items = (n >> 3) > (n & 7) ? (n >> 3) : (n & 7);
  const int grid = grid_for(items, threads, 8);
#define PS(M) \
  plms_step_kernel<T, M><<<grid, threads, 0, st>>>((const T*)x, (const T*)e0, (const T*)e1, \
                                                   (const T*)e2, (const T*)e3, sc, dA, denom, (T*)out, n)
  switch (mode) {
    case 0: PS(0); break;
    case 1: PS(1); break;
    case 2: PS(2); break;
    case 3: PS(3); break;
    default: PS(4); break;
  }
#undef PS
  return check_launch();
}


// ---------------------------------------------------------------------------
// a-2 + a-1 in ONE launch ("step_then_noise"): the reference runs the reverse update
// (segmentor.py:100-104) and the Laplace forward noising (ldiffusion.py:233-237) on latent-sized
// tensors inside the same loop iteration; both are elementwise over the same index space, and at the
// config shape (2 MiB) each stand-alone launch is pure latency.  Per 8 elements:
//   prev  = sc * x - (dA * eh) / denom            (exactly plms_step_kernel's chain)
//   noisy = clean + Laplace(0, b)                  (exactly laplace_qsample_kernel's chain)
// Same helper functions as the two kernels above, so every output bit is the same.
// ---------------------------------------------------------------------------
template <typename T, int MODE, int SRC>
__global__ void __launch_bounds__(256)
plms_step_noise_kernel(const T* __restrict__ x, const T* __restrict__ e0, const T* __restrict__ e1,
                       const T* __restrict__ e2, const T* __restrict__ e3, float sc, float dA, float denom,
                       T* __restrict__ prev, const T* __restrict__ clean, const T* __restrict__ inj,
                       T* __restrict__ noisy, float b, uint2 key, uint64_t offset, int64_t n) {
  const int64_t nvec = n >> 3;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
    const int64_t base = v << 3;
    float xv[8], a[8], bb[8], c[8], d[8], cl[8], nz[8];
    Vec8<T>::load(x + base, xv);
    Vec8<T>::load(e0 + base, a);
    if (MODE >= 1) Vec8<T>::load(e1 + base, bb);
    if (MODE >= 3) Vec8<T>::load(e2 + base, c);
    if (MODE >= 4) Vec8<T>::load(e3 + base, d);
    Vec8<T>::load(clean + base, cl);
    if (FusedAdd<T, SRC>::value) {                     // (exactly laplace_qsample_kernel's non-emitting form)
      float lg[8];
      uint32_t sg[8];
      philox_half8<Rounds<SRC>::value>(base, key, offset, lg, sg);
      const uint32_t kb = __float_as_uint(b * -0.6931471805599453f);
#pragma unroll
      for (int i = 0; i < 8; ++i) cl[i] = __fmaf_rn(lg[i], __uint_as_float(kb ^ sg[i]), cl[i]);
    } else {
      laplace_noise8<T, SRC>(inj, base, b, key, offset, nz);
#pragma unroll
      for (int i = 0; i < 8; ++i) cl[i] = __fadd_rn(cl[i], nz[i]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float eh = plms_eps<MODE>(a[i], MODE >= 1 ? bb[i] : 0.f, MODE >= 3 ? c[i] : 0.f,
                                      MODE >= 4 ? d[i] : 0.f);
      xv[i] = __fsub_rn(__fmul_rn(sc, xv[i]), __fdiv_rn(__fmul_rn(dA, eh), denom));
    }
    Vec8<T>::store(prev + base, xv);
    Vec8<T>::store(noisy + base, cl);
  }
  const int64_t t = (nvec << 3) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) {
    const float eh = plms_eps<MODE>(to_f32(e0[t]), MODE >= 1 ? to_f32(e1[t]) : 0.f,
                                    MODE >= 3 ? to_f32(e2[t]) : 0.f, MODE >= 4 ? to_f32(e3[t]) : 0.f);
    prev[t] = from_f32<T>(__fsub_rn(__fmul_rn(sc, to_f32(x[t])), __fdiv_rn(__fmul_rn(dA, eh), denom)));
    noisy[t] = from_f32<T>(__fadd_rn(to_f32(clean[t]), laplace_noise1<T, SRC>(inj, t, b, key, offset)));
  }
}

template <typename T>
static int launch_plms_noise(const void* x, const void* e0, const void* e1, const void* e2, const void* e3,
                             int mode, float sc, float dA, float denom, void* prev, const void* clean,
                             const void* noise_in, const void* u_in, void* noisy, float b, uint64_t seed,
                             uint64_t offset, int64_t n, cudaStream_t st) {
  const int threads = 256;
  const int64_t items = (n >> 3) > (n & 7) ? (n >> 3) : (n & 7);
  const int grid = grid_for(items, threads, 8);
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
#define PSN(M, SRC, INJ)                                                                               \
  plms_step_noise_kernel<T, M, SRC><<<grid, threads, 0, st>>>(                                         \
      (const T*)x, (const T*)e0, (const T*)e1, (const T*)e2, (const T*)e3, sc, dA, denom, (T*)prev,    \
      (const T*)clean, (const T*)(INJ), (T*)noisy, b, key, offset, n)
#define PSM(M)                                          \
  do {                                                  \
    if (noise_in) PSN(M, SRC_NOISE, noise_in);          \
    else if (u_in) PSN(M, SRC_UNIFORM, u_in);           \
    else if (philox_seven<T>()) PSN(M, SRC_PHILOX7, nullptr); \
    else PSN(M, SRC_PHILOX, nullptr);                   \
  } while (0)
  switch (mode) {
    case 0: PSM(0); break;
    case 1: PSM(1); break;
    case 2: PSM(2); break;
    case 3: PSM(3); break;
    default: PSM(4); break;
  }
#undef PSM
#undef PSN
  return check_launch();
}


// ---------------------------------------------------------------------------
// a-1 variant (segmentor.py:344-345 / :375): Laplace(0,1) noise modulated by a per-pixel
// scale map, and its inverse.
//   noising : out = fl(x_mul * x) + fl(noise * s)          noise ~ Laplace(0, 1)
//   inverse : out = fl(fl(x - fl(eps * s)) / out_div)
// The map is either element-for-element ([B,C,plane], span == 0) or one plane per image
// broadcast over the C channels ([B,1,plane], span == C*plane: what the reference
// materialises with .repeat(1, C, 1, 1)).  Vector path: 8 elements per thread; with a
// broadcast map it needs plane % 8 == 0 so that a vector never straddles two planes.
// Everything the vector path does not cover is done by a scalar grid-stride loop.
// ---------------------------------------------------------------------------
// index into the scale map of element i.  span == 0: the map has the tensor's shape.  Otherwise the
// map is [B, 1, plane] under a [B, C, plane] tensor (span = C * plane): image i / span, pixel i % plane.
// Power-of-two planes and channel counts (32x32x4, 128x128x4 latents) take shifts; other shapes one
// 32-bit division pair while the tensor has fewer than 2^31 elements (a 64-bit pair costs ~25
// instructions per element of an 8-wide vector and made the kernel ALU-bound at 0.85 of the roofline).
struct MapIdx {
  int64_t plane, span;
  int sh_span, sh_plane;          // >= 0: both are powers of two
  int small;                      // n < 2^31: 32-bit arithmetic is exact
};
__device__ __forceinline__ int64_t map_index(int64_t i, const MapIdx& m) {
  if (m.span == 0) return i;
  if (m.sh_span >= 0) return ((i >> m.sh_span) << m.sh_plane) | (i & (m.plane - 1));
  if (m.small) {
    const uint32_t u = (uint32_t)i, sp = (uint32_t)m.span, pl = (uint32_t)m.plane;
    return (int64_t)((u / sp) * pl + (u % pl));
  }
  return (i / m.span) * m.plane + (i % m.plane);
}

template <typename T, int SRC, bool EMIT>
__global__ void __launch_bounds__(256)
laplace_qsample_map_kernel(const T* __restrict__ x, const T* __restrict__ scale, T* __restrict__ out,
                           const T* __restrict__ inj, T* __restrict__ noise_out, float x_mul, uint2 key,
                           uint64_t offset, int64_t n, MapIdx mi, int64_t nvec) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t v = tid; v < nvec; v += stride) {
    const int64_t base = v << 3;
    float xv[8], nz[8], sv[8];
    Vec8<T>::load(x + base, xv);
    Vec8<T>::load(scale + map_index(base, mi), sv);
    laplace_noise8<T, SRC>(inj, base, 1.f, key, offset, nz);
    if (EMIT) Vec8<T>::store(noise_out + base, nz);
#pragma unroll
    for (int i = 0; i < 8; ++i) xv[i] = __fadd_rn(__fmul_rn(x_mul, xv[i]), __fmul_rn(nz[i], sv[i]));
    Vec8<T>::store(out + base, xv);
  }
  for (int64_t t = (nvec << 3) + tid; t < n; t += stride) {
    const float nzs = laplace_noise1<T, SRC>(inj, t, 1.f, key, offset);
    if (EMIT) noise_out[t] = from_f32<T>(nzs);
    const float s = to_f32(scale[map_index(t, mi)]);
    out[t] = from_f32<T>(__fadd_rn(__fmul_rn(x_mul, to_f32(x[t])), __fmul_rn(nzs, s)));
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
scaled_residual_kernel(const T* __restrict__ x, const T* __restrict__ eps, const T* __restrict__ scale,
                       T* __restrict__ out, float out_div, int64_t n, MapIdx mi, int64_t nvec) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t v = tid; v < nvec; v += stride) {
    const int64_t base = v << 3;
    float xv[8], ev[8], sv[8];
    Vec8<T>::load(x + base, xv);
    Vec8<T>::load(eps + base, ev);
    Vec8<T>::load(scale + map_index(base, mi), sv);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      xv[i] = __fdiv_rn(__fsub_rn(xv[i], __fmul_rn(ev[i], sv[i])), out_div);
    Vec8<T>::store(out + base, xv);
  }
  for (int64_t t = (nvec << 3) + tid; t < n; t += stride) {
    const float s = to_f32(scale[map_index(t, mi)]);
    out[t] = from_f32<T>(__fdiv_rn(__fsub_rn(to_f32(x[t]), __fmul_rn(to_f32(eps[t]), s)), out_div));
  }
}

// vector count of the map kernels: 0 when a broadcast plane is not a multiple of the vector width
static int64_t map_nvec(int64_t n, int64_t plane, int64_t span) {
  return (span != 0 && (plane & 7)) ? 0 : (n >> 3);
}
static int log2_exact(int64_t v) {                      // log2(v) if v is a power of two, else -1
  if (v <= 0 || (v & (v - 1))) return -1;
  int s = 0;
  while ((int64_t(1) << s) != v) ++s;
  return s;
}
static MapIdx map_idx(int64_t n, int64_t plane, int64_t span) {
  MapIdx m;
  m.plane = plane; m.span = span;
  m.sh_plane = log2_exact(plane);
  m.sh_span = (span != 0 && m.sh_plane >= 0) ? log2_exact(span) : -1;
  m.small = n < (int64_t(1) << 31);
  return m;
}

template <typename T>
static int launch_qsample_map(const void* x, const void* scale, void* out, const void* noise_in,
                              const void* u_in, void* noise_out, float x_mul, uint64_t seed,
                              uint64_t offset, int64_t n, int64_t plane, int64_t span, cudaStream_t st) {
  const int threads = 256;
  const int64_t nvec = map_nvec(n, plane, span), rest = n - (nvec << 3);
  const int grid = grid_for(nvec > rest ? nvec : rest, threads, 8);
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  T* no = (T*)noise_out;
#define LQM(SRC, EMIT, INJ)                                                                          \
  laplace_qsample_map_kernel<T, SRC, EMIT><<<grid, threads, 0, st>>>(                                \
      (const T*)x, (const T*)scale, (T*)out, (const T*)(INJ), no, x_mul, key, offset, n, map_idx(n, plane, span), nvec)
  if (noise_in) {
    if (no) LQM(SRC_NOISE, true, noise_in); else LQM(SRC_NOISE, false, noise_in);
  } else if (u_in) {
    if (no) LQM(SRC_UNIFORM, true, u_in); else LQM(SRC_UNIFORM, false, u_in);
  } else {
    if (philox_seven<T>()) { if (no) LQM(SRC_PHILOX7, true, nullptr); else LQM(SRC_PHILOX7, false, nullptr); }
    else if (no) LQM(SRC_PHILOX, true, nullptr); else LQM(SRC_PHILOX, false, nullptr);
  }
#undef LQM
  return check_launch();
}

template <typename T>
static int launch_scaled_residual(const void* x, const void* eps, const void* scale, void* out,
                                  float out_div, int64_t n, int64_t plane, int64_t span, cudaStream_t st) {
  const int64_t nvec = map_nvec(n, plane, span), rest = n - (nvec << 3);
  const int grid = grid_for(nvec > rest ? nvec : rest, 256, 8);
  scaled_residual_kernel<T><<<grid, 256, 0, st>>>((const T*)x, (const T*)eps, (const T*)scale, (T*)out,
                                                  out_div, n, map_idx(n, plane, span), nvec);
  return check_launch();
}

// shared argument check of the two map entry points; *span = 0 (full map) or C*plane (broadcast)
static int map_args(int64_t n, int64_t plane, int channels, int scale_channels, int64_t* span) {
  if (n < 0 || plane <= 0 || channels <= 0) return LDIFF_EINVAL;
  if (scale_channels != 1 && scale_channels != channels) return LDIFF_EINVAL;
  if (n % (plane * (int64_t)channels)) return LDIFF_EINVAL;
  *span = (scale_channels == channels) ? 0 : plane * (int64_t)channels;
  return LDIFF_OK;
}

}  // namespace ldiff

using namespace ldiff;

extern "C" int ldiff_laplace_qsample(const void* x, void* out, const void* noise_in, const void* u_in,
                                     void* noise_out, float b, uint64_t seed, uint64_t offset,
                                     int64_t n, int dtype, void* stream) {
  if (!x || !out || n < 0 || (noise_in && u_in)) return LDIFF_EINVAL;
  if (n == 0) return LDIFF_OK;
  if (!aligned16(x) || !aligned16(out) || !aligned16(noise_in) || !aligned16(u_in) ||
      !aligned16(noise_out))
    return LDIFF_EALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == LDIFF_F32)
    return launch_qsample<float>(x, out, noise_in, u_in, noise_out, b, seed, offset, n, st);
  if (dtype == LDIFF_BF16)
    return launch_qsample<__nv_bfloat16>(x, out, noise_in, u_in, noise_out, b, seed, offset, n, st);
  return LDIFF_EUNSUPPORTED;
}

extern "C" int ldiff_plms_step(const void* sample, const void* e0, const void* e1, const void* e2,
                               const void* e3, int mode, float sample_coeff, float alpha_diff,
                               float denom, void* prev_sample, int64_t n, int dtype, void* stream) {
  if (!sample || !e0 || !prev_sample || n < 0 || mode < 0 || mode > 4) return LDIFF_EINVAL;
  if ((mode >= 1 && !e1) || (mode >= 3 && !e2) || (mode >= 4 && !e3)) return LDIFF_EINVAL;
  if (n == 0) return LDIFF_OK;
  if (!aligned16(sample) || !aligned16(e0) || !aligned16(e1) || !aligned16(e2) || !aligned16(e3) ||
      !aligned16(prev_sample))
    return LDIFF_EALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == LDIFF_F32)
    return launch_plms<float>(sample, e0, e1, e2, e3, mode, sample_coeff, alpha_diff, denom,
                              prev_sample, n, st);
  if (dtype == LDIFF_BF16)
    return launch_plms<__nv_bfloat16>(sample, e0, e1, e2, e3, mode, sample_coeff, alpha_diff, denom,
                                      prev_sample, n, st);
  return LDIFF_EUNSUPPORTED;
}

extern "C" int ldiff_plms_step_noise(const void* sample, const void* e0, const void* e1, const void* e2,
                                     const void* e3, int mode, float sample_coeff, float alpha_diff,
                                     float denom, void* prev_sample, const void* clean, void* noisy,
                                     const void* noise_in, const void* u_in, float b, uint64_t seed,
                                     uint64_t offset, int64_t n, int dtype, void* stream) {
  if (!sample || !e0 || !prev_sample || !clean || !noisy || n < 0 || mode < 0 || mode > 4 || (noise_in && u_in))
    return LDIFF_EINVAL;
  if ((mode >= 1 && !e1) || (mode >= 3 && !e2) || (mode >= 4 && !e3)) return LDIFF_EINVAL;
  if (n == 0) return LDIFF_OK;
  if (!aligned16(sample) || !aligned16(e0) || !aligned16(e1) || !aligned16(e2) || !aligned16(e3) ||
      !aligned16(prev_sample) || !aligned16(clean) || !aligned16(noisy) || !aligned16(noise_in) || !aligned16(u_in))
    return LDIFF_EALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == LDIFF_F32)
    return launch_plms_noise<float>(sample, e0, e1, e2, e3, mode, sample_coeff, alpha_diff, denom, prev_sample,
                                    clean, noise_in, u_in, noisy, b, seed, offset, n, st);
  if (dtype == LDIFF_BF16)
    return launch_plms_noise<__nv_bfloat16>(sample, e0, e1, e2, e3, mode, sample_coeff, alpha_diff, denom,
                                            prev_sample, clean, noise_in, u_in, noisy, b, seed, offset, n, st);
  return LDIFF_EUNSUPPORTED;
}

extern "C" int ldiff_laplace_qsample_map(const void* x, const void* scale, void* out, const void* noise_in,
                                         const void* u_in, void* noise_out, float x_mul, uint64_t seed,
                                         uint64_t offset, int64_t n, int64_t plane, int channels,
                                         int scale_channels, int dtype, void* stream) {
  if (!x || !scale || !out || (noise_in && u_in)) return LDIFF_EINVAL;
  int64_t span = 0;
  const int rc = map_args(n, plane, channels, scale_channels, &span);
  if (rc != LDIFF_OK) return rc;
  if (n == 0) return LDIFF_OK;
  if (!aligned16(x) || !aligned16(scale) || !aligned16(out) || !aligned16(noise_in) || !aligned16(u_in) ||
      !aligned16(noise_out))
    return LDIFF_EALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == LDIFF_F32)
    return launch_qsample_map<float>(x, scale, out, noise_in, u_in, noise_out, x_mul, seed, offset, n, plane,
                                     span, st);
  if (dtype == LDIFF_BF16)
    return launch_qsample_map<__nv_bfloat16>(x, scale, out, noise_in, u_in, noise_out, x_mul, seed, offset, n,
                                             plane, span, st);
  return LDIFF_EUNSUPPORTED;
}

extern "C" int ldiff_scaled_residual(const void* x, const void* eps, const void* scale, void* out,
                                     float out_div, int64_t n, int64_t plane, int channels,
                                     int scale_channels, int dtype, void* stream) {
  if (!x || !eps || !scale || !out || !(out_div != 0.f)) return LDIFF_EINVAL;
  int64_t span = 0;
  const int rc = map_args(n, plane, channels, scale_channels, &span);
  if (rc != LDIFF_OK) return rc;
  if (n == 0) return LDIFF_OK;
  if (!aligned16(x) || !aligned16(eps) || !aligned16(scale) || !aligned16(out)) return LDIFF_EALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == LDIFF_F32)
    return launch_scaled_residual<float>(x, eps, scale, out, out_div, n, plane, span, st);
  if (dtype == LDIFF_BF16)
    return launch_scaled_residual<__nv_bfloat16>(x, eps, scale, out, out_div, n, plane, span, st);
  return LDIFF_EUNSUPPORTED;
}
