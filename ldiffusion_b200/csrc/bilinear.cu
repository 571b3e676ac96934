// a-4 bilinear lift (align_corners=False, no antialias) + weighted gray + channel
// concat into a preallocated [B,Ctot,H,W] buffer (sm_100a).
//
// Arithmetic is ATen's upsample_bilinear2d with every rounding pinned to what
// the CPU build of the reference's torch computes (oracle/bilinear.py):
//   src = fma(scale, dst + 0.5, -0.5), scale = fl(in/out); clamp at 0
//   t   = fma(w0, a, fl(w1*b))            horizontal
//   v   = fma(h0, t_top, fl(h1*t_bot))    vertical
//   gray = fl(fl(fl(0.2989 R) + fl(0.5870 G)) + fl(0.1140 B))
//
// Two kernels:
//  * lift_gather_kernel  — any sizes; one thread per output pixel gathers its
//    2x2 footprint (the 16x down-sample of the training path touches 4/256 of
//    the source; output is tiny).
//  * lift_separable_kernel — up-sampling: a CTA owns a band of output rows of
//    one image; the few source rows it needs are staged into shared memory with
//    1-D bulk TMA copies (cp.async.bulk + mbarrier), the horizontal pass is done
//    once per source row into REGISTERS, and every output row is then one mul +
//    one fma per pixel with 128-bit stores: HBM-write-bound.
#include <cstdlib>

#include "common.cuh"

namespace ldiff {

struct Axis {            // per-axis resampling constants
  float scale;           // fl(in / out)
  int in, out;
};

struct Tap { int i0, i1; float l0, l1; };

__device__ __forceinline__ Tap make_tap(const Axis& a, int dst) {
  Tap t;
  if (a.in == a.out) {             // ATen copies when the scale is 1
    t.i0 = t.i1 = dst; t.l0 = 1.f; t.l1 = 0.f;
    return t;
  }
  float src = __fmaf_rn(a.scale, (float)dst + 0.5f, -0.5f);
  src = fmaxf(src, 0.f);
  t.i0 = min((int)floorf(src), a.in - 1);
  t.i1 = min(t.i0 + 1, a.in - 1);
  t.l1 = fminf(fmaxf(__fsub_rn(src, (float)t.i0), 0.f), 1.f);
  t.l0 = __fsub_rn(1.f, t.l1);
  return t;
}

__device__ __forceinline__ float lerp_h(float w0, float a, float w1, float b) {
  return __fmaf_rn(w0, a, __fmul_rn(w1, b));
}

__device__ __forceinline__ float gray3(float r, float g, float b) {
  return __fadd_rn(__fadd_rn(__fmul_rn(0.2989f, r), __fmul_rn(0.5870f, g)), __fmul_rn(0.1140f, b));
}

// ----------------------------------------------------------------------------
constexpr int kMaxSrc = 8;
struct SrcList { const void* p[kMaxSrc]; };   // same-shaped sources of one launch (blockIdx.y picks one)

template <typename TS, typename TD>
__global__ void __launch_bounds__(256)
lift_gather_kernel(SrcList srcs, int C, int64_t sbs, int64_t scs, TD* __restrict__ dst,
                   int Ctot, int dch0, Axis ay, Axis ax, int B, int gray) {
  const TS* __restrict__ src = static_cast<const TS*>(srcs.p[blockIdx.y]);
  const int dch = dch0 + (int)blockIdx.y * (gray ? 1 : C);
  const int64_t HW = (int64_t)ay.out * ax.out;
  const int64_t total = HW * B;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int b = (int)(i / HW);
    const int64_t p = i - (int64_t)b * HW;
    const int y = (int)(p / ax.out), x = (int)(p - (int64_t)y * ax.out);
    const Tap ty = make_tap(ay, y), tx = make_tap(ax, x);
    const TS* sb = src + b * sbs;
    float v[3];
    const int nc = gray ? 3 : C;
    for (int c0 = 0; c0 < nc; c0 += 3) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int c = c0 + k;
        if (c < nc) {
          const TS* sc = sb + c * scs;
          const TS* r0 = sc + (int64_t)ty.i0 * ax.in;
          const TS* r1 = sc + (int64_t)ty.i1 * ax.in;
          const float top = lerp_h(tx.l0, to_f32(__ldg(r0 + tx.i0)), tx.l1, to_f32(__ldg(r0 + tx.i1)));
          const float bot = lerp_h(tx.l0, to_f32(__ldg(r1 + tx.i0)), tx.l1, to_f32(__ldg(r1 + tx.i1)));
          v[k] = lerp_h(ty.l0, top, ty.l1, bot);
          if (!gray) dst[((int64_t)b * Ctot + dch + c) * HW + p] = from_f32<TD>(v[k]);
        }
      }
    }
    if (gray) dst[((int64_t)b * Ctot + dch) * HW + p] = from_f32<TD>(gray3(v[0], v[1], v[2]));
  }
}

// ----------------------------------------------------------------------------
// mbarrier / bulk-TMA helpers (PTX; SASS: SYNCS.*, UBLKCP)
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra LAB_DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "LAB_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

template <int V>
__device__ __forceinline__ void load_row(const float* p, float* v) {     // p is 16-byte aligned
#pragma unroll
  for (int j = 0; j < V / 4; ++j) {
    const float4 q = reinterpret_cast<const float4*>(p)[j];
    v[4 * j] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w;
  }
}

template <typename TD> struct OutVec;   // 16 bytes of output per thread
template <> struct OutVec<float> {
  static constexpr int N = 4;
  __device__ static void store(float* p, const float* v) {
    __stcs(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
  }
};
template <> struct OutVec<__nv_bfloat16> {
  static constexpr int N = 8;
  __device__ static void store(__nv_bfloat16* p, const float* v) {
    Vec8<__nv_bfloat16>::store(p, *reinterpret_cast<const float(*)[8]>(v));
  }
};
template <> struct OutVec<uint8_t> {
  static constexpr int N = 16;
  __device__ static void store(uint8_t* p, const float* v) {
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      w[j] = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) w[j] |= (uint32_t)from_f32<uint8_t>(v[4 * j + k]) << (8 * k);
    }
    __stcs(reinterpret_cast<uint4*>(p), make_uint4(w[0], w[1], w[2], w[3]));
  }
};

// grid: (bands, channel groups, B).  NC = channels handled by one CTA (3 in gray
// mode, else 1).  Shared memory: [NC][max_rows][w] TS raw source rows (the bulk-TMA
// destination) + the band's vertical taps.  A thread owns V adjacent output
// columns (one 16-byte store) and walks down its rows keeping the horizontally
// lifted values of the two current source rows in REGISTERS; they are refreshed
// from the staged raw rows only when the source row pair changes (every H/h output
// rows), so the steady state per output element is one FMUL + one FFMA + the store.
template <typename TS, typename TD, int NC, bool GRAY>
__global__ void __launch_bounds__(256)
lift_separable_kernel(const TS* __restrict__ src, int64_t sbs, int64_t scs, TD* __restrict__ dst,
                      int Ctot, int dch, Axis ay, Axis ax, int band, int max_rows) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ int4 s_tap[64];                           // {r0, r1, bits(l0), bits(l1)} per band row
  const int w = ax.in, W = ax.out, H = ay.out;
  TS* raw = reinterpret_cast<TS*>(smem);

  const int b = blockIdx.z;
  const int c0 = blockIdx.y * NC;
  const int Y0 = blockIdx.x * band;
  const int Y1 = min(Y0 + band, H);
  const int ylo = make_tap(ay, Y0).i0;
  const int yhi = make_tap(ay, Y1 - 1).i1;
  const int nrows = yhi - ylo + 1;                      // <= max_rows by construction

  for (int i = threadIdx.x; i < Y1 - Y0; i += blockDim.x) {   // (blockDim.x can be smaller than the band)
    const Tap t = make_tap(ay, Y0 + i);
    s_tap[i] = make_int4(t.i0 - ylo, t.i1 - ylo, __float_as_int(t.l0), __float_as_int(t.l1));
  }
  if (threadIdx.x == 0) mbar_init(&bar, 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t row_bytes = (uint32_t)(w * sizeof(TS));
    mbar_expect_tx(&bar, row_bytes * nrows * NC);
    for (int c = 0; c < NC; ++c) {
      // the nrows source rows of one channel are contiguous in global memory
      bulk_g2s(raw + (size_t)c * max_rows * w, src + b * sbs + (c0 + c) * scs + (int64_t)ylo * w,
               row_bytes * nrows, &bar);
    }
  }
  mbar_wait(&bar, 0);

  constexpr int V = OutVec<TD>::N;
  const int groups = W / V;                             // 16-byte vectors per output row
  const int nthr = blockDim.x;
  // groups < nthr: the block splits the band into nsub row sub-bands
  const int nsub = groups >= nthr ? 1 : nthr / groups;
  const int sub = groups >= nthr ? 0 : (int)threadIdx.x / groups;
  const int xg0 = groups >= nthr ? (int)threadIdx.x : (int)threadIdx.x - sub * groups;
  const int xstep = groups >= nthr ? nthr : groups;
  if (sub >= nsub) return;
  const int rows_total = Y1 - Y0;
  const int rows_per_sub = (rows_total + nsub - 1) / nsub;
  const int ya = sub * rows_per_sub, yb = min(ya + rows_per_sub, rows_total);
  const int64_t HW = (int64_t)H * W;

  for (int xg = xg0; xg < groups; xg += xstep) {
    const int X = xg * V;
    float T0[NC][V], T1[NC][V];
    int xi0[V];                                         // horizontal taps of this thread's columns
    float xl1[V];
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const Tap tx = make_tap(ax, X + k);
      xi0[k] = tx.i0; xl1[k] = tx.l1;
    }
    auto fill = [&](float (&T)[NC][V], int r) {
#pragma unroll
      for (int k = 0; k < V; ++k) {
        const int i1 = (ax.in == ax.out) ? xi0[k] : min(xi0[k] + 1, w - 1);   // as make_tap
        const float l0 = __fsub_rn(1.f, xl1[k]);
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const TS* row = raw + ((size_t)c * max_rows + r) * w;
          T[c][k] = lerp_h(l0, to_f32(row[xi0[k]]), xl1[k], to_f32(row[i1]));
        }
      }
    };
    int cur0 = -1, cur1 = -1;
    // output pointer of this thread's vector in channel (dch [+ c0]) of the current row, advanced by W per row
    TD* orow = dst + ((int64_t)b * Ctot + dch + (GRAY ? 0 : c0)) * HW + (int64_t)(Y0 + ya) * W + X;
    for (int yy = ya; yy < yb; ++yy, orow += W) {
      const int4 tp = s_tap[yy];
      const float l0 = __int_as_float(tp.z), l1 = __int_as_float(tp.w);
      if (tp.x != cur0 || tp.y != cur1) {                // block-uniform per sub-band
        if (tp.x == cur1) {
#pragma unroll
          for (int c = 0; c < NC; ++c)
#pragma unroll
            for (int k = 0; k < V; ++k) T0[c][k] = T1[c][k];
        } else {
          fill(T0, tp.x);
        }
        if (tp.y == tp.x) {
#pragma unroll
          for (int c = 0; c < NC; ++c)
#pragma unroll
            for (int k = 0; k < V; ++k) T1[c][k] = T0[c][k];
        } else {
          fill(T1, tp.y);
        }
        cur0 = tp.x; cur1 = tp.y;
      }
      float out[V];
      if (GRAY) {
        float ch[3][V];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int k = 0; k < V; ++k) ch[c][k] = lerp_h(l0, T0[c % NC][k], l1, T1[c % NC][k]);
#pragma unroll
        for (int k = 0; k < V; ++k) out[k] = gray3(ch[0][k], ch[1][k], ch[2][k]);
        OutVec<TD>::store(orow, out);
      } else {
        // two elements per instruction on the packed fp32x2 pipe: FMUL2 + FFMA2 are IEEE per lane, i.e. exactly
        // lerp_h's fma(l0, a, fl(l1 * b)) (the product feeds the ADDEND of the fma, so nothing can be contracted)
        const float2 l0p = make_float2(l0, l0), l1p = make_float2(l1, l1);
#pragma unroll
        for (int c = 0; c < NC; ++c) {
#pragma unroll
          for (int k = 0; k < V; k += 2) {
            const float2 v = __ffma2_rn(l0p, make_float2(T0[c][k], T0[c][k + 1]),
                                        __fmul2_rn(l1p, make_float2(T1[c][k], T1[c][k + 1])));
            out[k] = v.x; out[k + 1] = v.y;
          }
          OutVec<TD>::store(orow + c * HW, out);
        }
      }
    }
  }
}

template <typename TS, typename TD>
static int launch_lift(const void* src, int C, int h, int w, int64_t sbs, int64_t scs, void* dst,
                       int Ctot, int dch, int H, int W, int B, int gray, cudaStream_t st) {
  Axis ay{(float)h / (float)H, h, H}, ax{(float)w / (float)W, w, W};
  const int threads = 256;
  constexpr int V = OutVec<TD>::N;
  static const int band_knob = [] { const char* e = getenv("LDIFF_LIFT_BAND"); return e ? atoi(e) : 0; }();
  // measured at 64 -> 1024: 32-row bands are best for 4-byte outputs (16.6 vs 18.4 us), 64-row bands
  // for 2-byte outputs (10.9 vs 11.3 us: half the bytes per band, same start-up cost)
  const int band = (band_knob == 32 || band_knob == 64) ? band_knob : (sizeof(TD) <= 2 ? 64 : 32);
  // rows of source a band can touch: ceil(band * scale) + 2
  const int max_rows = (int)((double)band * h / H) + 3;
  const int NCg = gray ? 3 : 1;
  const size_t smem = ((size_t)NCg * max_rows * w * sizeof(TS) + 127) & ~(size_t)127;
  const bool separable = H >= h && W >= w && (W % V == 0) && ((w * sizeof(TS)) % 16 == 0) &&
                         aligned16(src) && aligned16(dst) && ((sbs * sizeof(TS)) % 16 == 0) &&
                         ((scs * sizeof(TS)) % 16 == 0) && (((int64_t)H * W * sizeof(TD)) % 16 == 0) &&
                         smem <= 160 * 1024 && (int64_t)max_rows * w * sizeof(TS) < (1 << 20);
  if (separable) {
    dim3 grid((H + band - 1) / band, gray ? 1 : C, B);
    // one thread per 16-byte output vector of a row when the row has < 256 of them
    const int groups = W / V;
    const int threads = groups >= 256 ? 256 : (groups >= 128 ? 128 : (groups >= 64 ? 64 : 32));
    if (gray) {
      auto k = lift_separable_kernel<TS, TD, 3, true>;
      if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
      k<<<grid, threads, smem, st>>>((const TS*)src, sbs, scs, (TD*)dst, Ctot, dch, ay, ax, band, max_rows);
    } else {
      auto k = lift_separable_kernel<TS, TD, 1, false>;
      if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
      k<<<grid, threads, smem, st>>>((const TS*)src, sbs, scs, (TD*)dst, Ctot, dch, ay, ax, band, max_rows);
    }
  } else {
    const int64_t total = (int64_t)H * W * B;
    SrcList one{};
    one.p[0] = src;
    lift_gather_kernel<TS, TD><<<grid_for(total, threads, 8), threads, 0, st>>>(
        one, C, sbs, scs, (TD*)dst, Ctot, dch, ay, ax, B, gray);
  }
  return check_launch();
}

template <typename TS, typename TD>
static int launch_lift_multi(const void* const* srcs, int n_src, int C, int h, int w, int64_t sbs, int64_t scs,
                             void* dst, int Ctot, int dch, int H, int W, int B, int gray, cudaStream_t st) {
  Axis ay{(float)h / (float)H, h, H}, ax{(float)w / (float)W, w, W};
  SrcList list{};
  for (int i = 0; i < n_src; ++i) list.p[i] = srcs[i];
  const int64_t total = (int64_t)H * W * B;
  int bx = grid_for(total, 256, 8);
  const int cap = (sm_count() * 8 + n_src - 1) / n_src;
  if (bx > cap) bx = cap;
  lift_gather_kernel<TS, TD><<<dim3(bx, n_src), 256, 0, st>>>(list, C, sbs, scs, (TD*)dst, Ctot, dch, ay, ax, B,
                                                            gray);
  return check_launch();
}

}  // namespace ldiff

using namespace ldiff;

extern "C" int ldiff_bilinear_lift_multi(const void* const* host_srcs, int n_src, int src_dtype, int C, int h,
                                         int w, int64_t src_batch_stride, int64_t src_channel_stride, void* dst,
                                         int dst_dtype, int Ctot, int dst_channel, int H, int W, int B, int gray,
                                         void* stream) {
  if (!host_srcs || n_src < 1 || n_src > kMaxSrc || !dst || C < 1 || h < 1 || w < 1 || H < 1 || W < 1 || B < 0 ||
      dst_channel < 0)
    return LDIFF_EINVAL;
  for (int i = 0; i < n_src; ++i)
    if (!host_srcs[i]) return LDIFF_EINVAL;
  if (gray && C != 3) return LDIFF_EINVAL;
  if (dst_channel + n_src * (gray ? 1 : C) > Ctot) return LDIFF_EINVAL;
  if (B == 0) return LDIFF_OK;
  cudaStream_t st = (cudaStream_t)stream;
#define LIFTM(TS, TD)                                                                                  \
  return launch_lift_multi<TS, TD>(host_srcs, n_src, C, h, w, src_batch_stride, src_channel_stride, dst, \
                                   Ctot, dst_channel, H, W, B, gray, st)
  typedef __nv_bfloat16 bf16;
  if (src_dtype == LDIFF_F32 && dst_dtype == LDIFF_F32) LIFTM(float, float);
  if (src_dtype == LDIFF_BF16 && dst_dtype == LDIFF_BF16) LIFTM(bf16, bf16);
  if (src_dtype == LDIFF_BF16 && dst_dtype == LDIFF_F32) LIFTM(bf16, float);
  if (src_dtype == LDIFF_F32 && dst_dtype == LDIFF_BF16) LIFTM(float, bf16);
  if (src_dtype == LDIFF_U8 && dst_dtype == LDIFF_U8) LIFTM(uint8_t, uint8_t);
  if (src_dtype == LDIFF_U8 && dst_dtype == LDIFF_F32) LIFTM(uint8_t, float);
#undef LIFTM
  return LDIFF_EUNSUPPORTED;
}

extern "C" int ldiff_bilinear_lift(const void* src, int src_dtype, int C, int h, int w,
                                   int64_t src_batch_stride, int64_t src_channel_stride, void* dst,
                                   int dst_dtype, int Ctot, int dst_channel, int H, int W, int B,
                                   int gray, void* stream) {
  if (!src || !dst || C < 1 || h < 1 || w < 1 || H < 1 || W < 1 || B < 0 || dst_channel < 0)
    return LDIFF_EINVAL;
  if (gray && C != 3) return LDIFF_EINVAL;
  if (dst_channel + (gray ? 1 : C) > Ctot) return LDIFF_EINVAL;
  if (B == 0) return LDIFF_OK;
  cudaStream_t st = (cudaStream_t)stream;
#define LIFT(TS, TD)                                                                             \
  return launch_lift<TS, TD>(src, C, h, w, src_batch_stride, src_channel_stride, dst, Ctot,      \
                             dst_channel, H, W, B, gray, st)
  typedef __nv_bfloat16 bf16;
  if (src_dtype == LDIFF_F32 && dst_dtype == LDIFF_F32) LIFT(float, float);
  if (src_dtype == LDIFF_BF16 && dst_dtype == LDIFF_BF16) LIFT(bf16, bf16);
  if (src_dtype == LDIFF_BF16 && dst_dtype == LDIFF_F32) LIFT(bf16, float);
  if (src_dtype == LDIFF_F32 && dst_dtype == LDIFF_BF16) LIFT(float, bf16);
  if (src_dtype == LDIFF_U8 && dst_dtype == LDIFF_U8) LIFT(uint8_t, uint8_t);
  if (src_dtype == LDIFF_U8 && dst_dtype == LDIFF_F32) LIFT(uint8_t, float);
#undef LIFT
  return LDIFF_EUNSUPPORTED;
}
