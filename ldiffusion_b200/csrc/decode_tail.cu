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

// The kernel is issue-bound before it is HBM-bound unless the per-pixel
// instruction count stays near 15, so the chain is strength-reduced without
// changing a single rounding:
//  * x/2 is exact, so fl(x/2 + 0.5) == fma(x, 0.5, 0.5), and the clamp to [0,1]
//    is the FFMA's .SAT modifier (NaN -> 0, as numpy's uint8 cast gives);
//  * rint(v) for v in [0,255] == low mantissa bits of fl(v + 1.5*2^23)
//    (round-half-even in hardware, no F2I conversion);
//  * the 16.16 luma weights sum to exactly 2^16, so the 0x4B400000 exponent
//    pattern of those magic floats cancels mod 2^32: the luma is three IMADs on
//    the RAW float bits and its result byte is picked out with PRMT;
//  * bf16 images: the reference's bf16 VAE rounds x/2+0.5 to bf16 (x/2 is exact
//    there too): one packed fma.relu.bf16x2 + min.bf16x2 per TWO pixels, straight
//    on the loaded words; t then has 8 significant bits, t*255 is exact in fp32
//    and the multiply folds into the magic add as one FFMA.
template <typename T> struct Quant;

template <> struct Quant<float> {
  // returns the float bits 0x4B400000 | q
  __device__ static __forceinline__ uint32_t bits(float x) {
    const float t = __saturatef(__fmaf_rn(x, 0.5f, 0.5f));
    return __float_as_uint(__fadd_rn(__fmul_rn(t, 255.f), 12582912.f));
  }
};

template <> struct Quant<__nv_bfloat16> {
  // two packed bf16 pixels -> two magic-float bit patterns.  fma.rn.relu.bf16x2 rounds
  // x/2+0.5 to bf16 once; the reference rounds to fp32 first, which is exact here (x has 8
  // significant bits) or cannot change the bf16 result (|x| < 2^-16).  NaN inputs are
  // unspecified (numpy's uint8 cast of NaN is platform-defined too).
  __device__ static __forceinline__ void packed2(uint32_t x2, uint32_t& q0, uint32_t& q1) {
    uint32_t t;
    asm("{ .reg .b32 f;\n"
        "  fma.rn.relu.bf16x2 f, %1, %2, %2;\n"
        "  min.bf16x2 %0, f, %3;\n}"
        : "=r"(t) : "r"(x2), "r"(0x3f003f00u), "r"(0x3f803f80u));
    // both pixels in one packed fp32x2 FMA (FFMA2: IEEE per lane)
    const float2 y = __ffma2_rn(make_float2(__uint_as_float(t << 16), __uint_as_float(t & 0xffff0000u)),
                                make_float2(255.f, 255.f), make_float2(12582912.f, 12582912.f));
    q0 = __float_as_uint(y.x);
    q1 = __float_as_uint(y.y);
  }
  __device__ static __forceinline__ uint32_t bits(float x) {
    const __nv_bfloat16 h = __float2bfloat16_rn(x);            // x is a bf16 value: exact
    uint32_t a, b;
    packed2((uint32_t)*reinterpret_cast<const uint16_t*>(&h), a, b);
    return a;
  }
};

// 16 consecutive pixels of one plane -> 16 magic-float bit patterns
__device__ __forceinline__ void quant16(const float* p, uint32_t (&q)[16]) {
  float v[16];
  Vec8<float>::load(p, reinterpret_cast<float(&)[8]>(v[0]));
  Vec8<float>::load(p + 8, reinterpret_cast<float(&)[8]>(v[8]));
  // (scalar on purpose: ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2,
  //  which would round t*255 + magic once instead of twice)
#pragma unroll
  for (int i = 0; i < 16; ++i) q[i] = Quant<float>::bits(v[i]);
}
__device__ __forceinline__ void quant16(const __nv_bfloat16* p, uint32_t (&q)[16]) {
  const uint4 a = __ldcs(reinterpret_cast<const uint4*>(p));
  const uint4 b = __ldcs(reinterpret_cast<const uint4*>(p) + 1);
  const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) Quant<__nv_bfloat16>::packed2(w[i], q[2 * i], q[2 * i + 1]);
}

// (19595 R + 38470 G + 7471 B + 0x8000): result byte 2 is the PIL "L" value
__device__ __forceinline__ uint32_t luma_sum(uint32_t rb, uint32_t gb, uint32_t bb) {
  return 19595u * rb + 38470u * gb + 7471u * bb + 0x8000u;
}

// byte `B` of four registers -> one packed word
template <int B>
__device__ __forceinline__ uint32_t pack4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  const uint32_t lo = __byte_perm(a, b, 0x0040 + B * 0x11);
  const uint32_t hi = __byte_perm(c, d, 0x0040 + B * 0x11);
  return __byte_perm(lo, hi, 0x5410);
}

struct NormArgs {            // optional fused model-input output (N1 of SURVEY 8f)
  float* out;                // fp32 [B,3,H,W]: ((q/255) - mean) / std, torchvision ToTensor + Normalize
  float mean[3], stdv[3];
};

// Optional outputs fused into the tail (device pointers, null = off).  Valid for H % 16 == 0 and
// W % 16 == 0, where the reference's 16x bilinear down-sample (align_corners=False) is exactly the
// 2x2 footprint at rows / columns 16d+7, 16d+8 with weights 0.5: src = fma(16, d + 0.5, -0.5) = 16d + 7.5.
// The thread that owns columns 16g .. 16g+15 of row 16e+7 re-reads its two footprint pixels and the two
// below them (scalar loads, L1/L2 hits: 1/16 of the warps, 4 loads per channel) and evaluates the pinned
// lerp chain of bilinear.cu — the tail already streams every byte those gathers used to fetch again.
struct TailExtras {
  void* feat;                  // a-4 feature (ldiffusion.py:240-247): weighted gray of the lift -> channel feat_ch
  int feat_ctot, feat_ch;      //   of [B, feat_ctot, H/16, W/16]; storage = the image dtype, or fp32 (feat_f32)
  int feat_f32;
  void* small_rgb;             // [B,3,H/16,W/16] image dtype: the lift itself (source of ldiffusion.py:251)
  const uint8_t* label;        // uint8 [B,H,W] ground truth ...
  uint8_t* label_plane;        // ... copied into the label slot of the pixel vectors (pixel_latent_vector.py:92)
  int64_t label_plane_stride;
  uint8_t* label_small;        // ... and uint8 [B,1,H/16,W/16] = trunc(bilinear(label)) (ldiffusion.py:224-226)
  int W, fh, gpr, gpr_shift;   // row width, H/16, 16-pixel groups per row (= W/16), log2(gpr) or -1
};

__device__ __forceinline__ float half_lerp(float a, float b) {        // lerp_h(0.5, a, 0.5, b) of bilinear.cu
  return __fmaf_rn(0.5f, a, __fmul_rn(0.5f, b));
}

// The extras in two halves, so that their scattered loads are in flight while the thread does its main work:
// extras_load() issues them BEFORE the tile is read (nothing waits on the results yet), extras_finish()
// consumes them after the tile's stores.
struct ExtrasRegs {
  float px[3][4];              // footprint of each channel: top-left, top-right, bottom-left, bottom-right
  uint4 lab;                   // this thread's 16 label bytes (label plane copy)
  uint32_t lab4;               // the label's 2x2 footprint, one byte each
  int64_t o;                   // output pixel of a feature plane, -1: this thread owns no footprint
};

template <typename T>
__device__ __forceinline__ void extras_load(const T* __restrict__ img, const TailExtras& ex, int64_t b, int64_t hw,
                                            int gidx, ExtrasRegs& er) {
  const int64_t p = (int64_t)gidx << 4;
  if (ex.label_plane) er.lab = __ldg(reinterpret_cast<const uint4*>(ex.label + b * hw + p));
  const int y = ex.gpr_shift >= 0 ? (gidx >> ex.gpr_shift) : (gidx / ex.gpr);
  er.o = -1;
  if ((y & 15) != 7) return;
  er.o = ((int64_t)(y >> 4)) * ex.gpr + (gidx - y * ex.gpr);          // output pixel (y/16, g) of a plane
  if (ex.feat || ex.small_rgb) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const T* r0 = img + (b * 3 + c) * hw + p + 7;
      const T* r1 = r0 + ex.W;
      er.px[c][0] = to_f32(__ldg(r0)); er.px[c][1] = to_f32(__ldg(r0 + 1));
      er.px[c][2] = to_f32(__ldg(r1)); er.px[c][3] = to_f32(__ldg(r1 + 1));
    }
  }
  if (ex.label_small) {
    const uint8_t* r0 = ex.label + b * hw + p + 7;
    const uint8_t* r1 = r0 + ex.W;
    er.lab4 = (uint32_t)__ldg(r0) | ((uint32_t)__ldg(r0 + 1) << 8) | ((uint32_t)__ldg(r1) << 16) |
              ((uint32_t)__ldg(r1 + 1) << 24);
  }
}

template <typename T>
__device__ __forceinline__ void extras_finish(const TailExtras& ex, int64_t b, int gidx, const ExtrasRegs& er) {
  if (ex.label_plane)
    __stcs(reinterpret_cast<uint4*>(ex.label_plane + b * ex.label_plane_stride + ((int64_t)gidx << 4)), er.lab);
  if (er.o < 0) return;
  const int64_t fplane = (int64_t)ex.fh * ex.gpr;
  if (ex.feat || ex.small_rgb) {
    float v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      v[c] = half_lerp(half_lerp(er.px[c][0], er.px[c][1]), half_lerp(er.px[c][2], er.px[c][3]));
      if (ex.small_rgb) static_cast<T*>(ex.small_rgb)[(b * 3 + c) * fplane + er.o] = from_f32<T>(v[c]);
    }
    if (ex.feat) {
      const float gr = __fadd_rn(__fadd_rn(__fmul_rn(0.2989f, v[0]), __fmul_rn(0.5870f, v[1])), __fmul_rn(0.1140f, v[2]));
      const int64_t fo = (b * ex.feat_ctot + ex.feat_ch) * fplane + er.o;
      if (ex.feat_f32) static_cast<float*>(ex.feat)[fo] = gr;
      else static_cast<T*>(ex.feat)[fo] = from_f32<T>(gr);
    }
  }
  if (ex.label_small) {
    const float v = half_lerp(half_lerp((float)(er.lab4 & 0xffu), (float)((er.lab4 >> 8) & 0xffu)),
                              half_lerp((float)((er.lab4 >> 16) & 0xffu), (float)(er.lab4 >> 24)));
    ex.label_small[b * fplane + er.o] = from_f32<uint8_t>(v);
  }
}

template <typename T, bool RGB, bool GRAY, bool NORM, bool EXTRA = false>
__global__ void __launch_bounds__(256)
decode_tail_vec16_kernel(const T* __restrict__ img, uint8_t* __restrict__ rgb,
                         uint8_t* __restrict__ gray, int64_t hw, int groups_per_img,
                         int64_t gray_batch_stride, NormArgs na, TailExtras ex = TailExtras{}) {
  // a channel value has 256 possible inputs: the two IEEE divisions are tabulated once per block
  __shared__ float s_norm[NORM ? 3 * 256 : 1];
  if (NORM) {
    for (int i = threadIdx.x; i < 3 * 256; i += blockDim.x) {
      const int c = i >> 8, q = i & 255;
      s_norm[i] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)q, 255.f), na.mean[c]), na.stdv[c]);
    }
    __syncthreads();
  }
  // grid = (blocks over one image, B): no division in the loop
  const int64_t b = blockIdx.y;
  const int stride = gridDim.x * blockDim.x;
  for (int gidx = blockIdx.x * blockDim.x + threadIdx.x; gidx < groups_per_img; gidx += stride) {
    const int64_t p = (int64_t)gidx << 4;
    const T* base = img + b * 3 * hw + p;
    ExtrasRegs er;
    if (EXTRA) extras_load<T>(img, ex, b, hw, gidx, er);
    uint32_t q[3][16];                                 // float bits 0x4B400000 | value
#pragma unroll
    for (int c = 0; c < 3; ++c) quant16(base + c * hw, q[c]);
    if (GRAY) {
      uint32_t w[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        w[j] = pack4<2>(luma_sum(q[0][4 * j], q[1][4 * j], q[2][4 * j]),
                        luma_sum(q[0][4 * j + 1], q[1][4 * j + 1], q[2][4 * j + 1]),
                        luma_sum(q[0][4 * j + 2], q[1][4 * j + 2], q[2][4 * j + 2]),
                        luma_sum(q[0][4 * j + 3], q[1][4 * j + 3], q[2][4 * j + 3]));
      __stcs(reinterpret_cast<uint4*>(gray + b * gray_batch_stride + p),
             make_uint4(w[0], w[1], w[2], w[3]));
    }
    if (RGB) {
      uint32_t w[12];
#pragma unroll
      for (int j = 0; j < 12; ++j) {
        // output bytes 4j .. 4j+3: byte k belongs to pixel k/3, channel k%3
        w[j] = pack4<0>(q[(4 * j) % 3][(4 * j) / 3], q[(4 * j + 1) % 3][(4 * j + 1) / 3],
                        q[(4 * j + 2) % 3][(4 * j + 2) / 3], q[(4 * j + 3) % 3][(4 * j + 3) / 3]);
      }
      uint4* o = reinterpret_cast<uint4*>(rgb + (b * hw + p) * 3);
      __stcs(o, make_uint4(w[0], w[1], w[2], w[3]));
      __stcs(o + 1, make_uint4(w[4], w[5], w[6], w[7]));
      __stcs(o + 2, make_uint4(w[8], w[9], w[10], w[11]));
    }
    if (NORM) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float4* o = reinterpret_cast<float4*>(na.out + (b * 3 + c) * hw + p);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          __stcs(o + j, make_float4(s_norm[c * 256 + (q[c][4 * j] & 255)], s_norm[c * 256 + (q[c][4 * j + 1] & 255)],
                                    s_norm[c * 256 + (q[c][4 * j + 2] & 255)],
                                    s_norm[c * 256 + (q[c][4 * j + 3] & 255)]));
      }
    }
    if (EXTRA) extras_finish<T>(ex, b, gidx, er);
  }
}

// ----------------------------------------------------------------------------
// Bulk-TMA staged form: what the pass runs (bf16 default: 2 stages x 2 CTAs per SM; LDIFF_TUNE_DECODE_TAIL_TMA /
// LDIFF_DT_TMA pick another pipeline shape, 0 = the register-staged kernel above).  Why: the register-staged
// kernel's bytes in flight are tied to resident threads (6 blocks x 256 threads x 96 B), so inside the pass —
// where four other chains take SM time — it loses bandwidth in proportion to the issue slots it is denied
// (profiles/r01_pass_persist.txt).  Here a persistent CTA keeps STAGES tiles of 3 x kDtTile pixels in flight
// through 1-D bulk copies (cp.async.bulk + mbarrier complete_tx; no registers, no address arithmetic per
// load) and the threads only read shared memory.  Alone it is SLOWER than the register-staged kernel (11.1 vs
// 10.1 us for the gray plane of 8 x 1024^2 bf16: one CTA-wide hand-over per tile), inside the pass it is worth
// 10-17 us (profiles/r02_pass_time.txt).  Same arithmetic functions as above, so the bytes written are
// identical by construction.
// ----------------------------------------------------------------------------
constexpr int kDtTile = 4096;      // pixels per tile = 256 threads x 16 pixels
// a stage = 3 planes x 4096 px: 24 KB in bf16, 48 KB in fp32 (pipeline shape = template parameters STAGES x CTAS of
// the kernel; TmaShapes below lists the selectable ones)

__device__ __forceinline__ uint32_t dt_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void dt_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(dt_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void dt_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(dt_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void dt_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "DT_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DT_DONE;\n\t"
      "bra DT_WAIT;\n\t"
      "DT_DONE:\n\t}" ::"r"(dt_smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void dt_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(dt_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void dt_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   dt_smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(dt_smem_u32(bar))
               : "memory");
}

// 16 consecutive bf16 pixels of one plane, staged in shared memory -> 16 magic-float bit patterns
__device__ __forceinline__ void quant16_staged(const __nv_bfloat16* p, uint32_t (&q)[16]) {
  const uint4 a = *reinterpret_cast<const uint4*>(p);
  const uint4 b = *(reinterpret_cast<const uint4*>(p) + 1);
  const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) Quant<__nv_bfloat16>::packed2(w[i], q[2 * i], q[2 * i + 1]);
}

__device__ __forceinline__ void quant16_staged(const float* p, uint32_t (&q)[16]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 v = *(reinterpret_cast<const float4*>(p) + j);
    q[4 * j] = Quant<float>::bits(v.x);
    q[4 * j + 1] = Quant<float>::bits(v.y);
    q[4 * j + 2] = Quant<float>::bits(v.z);
    q[4 * j + 3] = Quant<float>::bits(v.w);
  }
}

// grid = persistent CTAs walking tiles of kDtTile pixels.  Without extras: tile t of the job = image
// t / tiles_per_img, pixels (t % tiles_per_img) * kDtTile ...  With extras the tiles of an image are SHIFTED by
// tile_off pixels (a whole number of rows; tiles_per_img then counts one more, the first and last tile are
// partial): the shift is chosen by the host so that the 2x2 footprints of the 16x down-sample (rows 16e+7, 16e+8)
// never straddle two tiles, and the thread that owns columns 16g..16g+15 of row 16e+7 reads all four footprint
// pixels of each channel from SHARED memory — the feature costs no global load at all.
// (Round-2 measurements at 8 x 1024^2 bf16, gray only, 2 stages x 3 CTAs: plain 11.3 us; footprints re-read
// from global by the owning threads 13.4 us; by a dedicated ninth warp 14.9 us (latency-bound, and the label
// copy through it 25 us); from shared memory 12.2 us with 2 x 2 — profiles/r02_kbench_fused.txt.)
// (register cap: 42 for the gray-only forms, 64 with RGB — the CTAs of this kernel are resident for a whole launch,
// and what they leave of the register file is what the concurrent chains of the pass get to run in)
// PW: a ninth warp is the PRODUCER (canonical TMA pipeline): it alone waits for a stage to be released (one
// mbarrier arrive per consumer warp) and refills it, so the eight consumer warps never meet at a CTA-wide barrier
// — each one goes from tile to tile as fast as its own data arrives.
// EXTRA: 0 = no fused extras, 1 = step feature / small RGB, 2 = those + the label outputs (a compile-time split: the
// label pointers' per-tile tests and predicated loads were ~20 issued instructions per warp and tile for nothing in the
// four tails of a pass that carry no label)
template <typename T, bool RGB, bool GRAY, int EXTRA, int STAGES, int CTAS, bool PW = false>
__global__ void __launch_bounds__(256 + (PW ? 32 : 0),
                                  (PW ? (RGB ? 3 : 5) : (RGB ? 4 : 6)) > CTAS ? (PW ? (RGB ? 3 : 5) : (RGB ? 4 : 6)) : CTAS)
decode_tail_tma_kernel(const T* __restrict__ img, uint8_t* __restrict__ rgb,
                       uint8_t* __restrict__ gray, int64_t hw, int tiles_per_img, int total_tiles,
                       int64_t gray_batch_stride, int tile_off, TailExtras ex) {
  constexpr int kDtStages = STAGES;
  extern __shared__ __align__(128) uint8_t dt_smem[];            // [kDtStages][3][kDtTile] T
  const int tiles_shift = (tiles_per_img & (tiles_per_img - 1)) == 0 ? 31 - __clz(tiles_per_img) : -1;
  __shared__ __align__(8) uint64_t full[kDtStages];
  __shared__ __align__(8) uint64_t empty[PW ? kDtStages : 1];
  T* buf = reinterpret_cast<T*>(dt_smem);
  const int tid = threadIdx.x;

  // tile (image b, index i in the image) -> first pixel p0, pixel count n (< kDtTile only for the partial tiles of a
  // shifted image).  (b, i) of a CTA's tile sequence t = blockIdx.x + k * gridDim.x is carried incrementally — one
  // division per CTA, not two per tile (257 tiles per image is not a power of two)
  auto advance = [&](int& b, int& i, int by) {
    i += by;
    while (i >= tiles_per_img) { i -= tiles_per_img; ++b; }
  };
  // (32-bit, branch-free: an image has fewer than 2^31 pixels — checked by the host; the EXTRA form spent 71 more
  //  instructions per warp and tile than the plain one on this bookkeeping, a quarter of the kernel)
  const int hw32 = (int)hw;
  const int shift_back = (EXTRA && tile_off) ? kDtTile - tile_off : 0;      // tile i starts at i * kDtTile - shift_back
  auto tile_of = [&](int i, int& p0, int& n) {
    if (!EXTRA) { p0 = i * kDtTile; n = kDtTile; return; }
    const int start = i * kDtTile - shift_back;
    p0 = max(start, 0);
    n = min(start + kDtTile, hw32) - p0;
  };
  // k-th tile of this CTA -> stage k % kDtStages (one thread issues; completion lands on full[stage])
  int cb = tiles_shift >= 0 ? ((int)blockIdx.x >> tiles_shift) : ((int)blockIdx.x / tiles_per_img);   // this CTA's first
  int ci = (int)blockIdx.x - cb * tiles_per_img;                                                       // tile
  auto issue = [&](int k, int b, int i) {                        // load the CTA's k-th tile = tile i of image b
    const int t = blockIdx.x + k * gridDim.x;
    if (t >= total_tiles) return;
    const int s = k % kDtStages;
    int n, p0;
    tile_of(i, p0, n);
    const uint32_t bytes = (uint32_t)n * sizeof(T);
    dt_mbar_expect_tx(&full[s], 3 * bytes);
#pragma unroll
    for (int c = 0; c < 3; ++c)
      dt_bulk_g2s(buf + (s * 3 + c) * kDtTile, img + ((int64_t)b * 3 + c) * hw + p0, bytes, &full[s]);
  };
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kDtStages; ++s) {
      dt_mbar_init(&full[s], 1);
      if (PW) dt_mbar_init(&empty[s], 8);                        // one arrive per consumer warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (PW && tid >= 256) {                                        // ---- producer warp (one lane issues)
    if (tid == 256) {
      for (int k = 0;; ++k) {
        if (blockIdx.x + k * gridDim.x >= total_tiles) break;
        if (k >= kDtStages) {
          dt_mbar_wait(&empty[k % kDtStages], (uint32_t)(k / kDtStages - 1) & 1u);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        issue(k, cb, ci);
        advance(cb, ci, (int)gridDim.x);
      }
    }
    return;
  }
  if (!PW && tid == 0) {
    int pb = cb, pi = ci;
#pragma unroll
    for (int k = 0; k < kDtStages; ++k) {
      issue(k, pb, pi);
      advance(pb, pi, (int)gridDim.x);
    }
  }
  for (int k = 0;; ++k) {
    const int t = blockIdx.x + k * gridDim.x;
    if (t >= total_tiles) break;                                 // (block-uniform)
    const int s = k % kDtStages;
    int n, p0;
    tile_of(ci, p0, n);
    const int64_t b = cb;
    advance(cb, ci, (int)gridDim.x);
    const bool mine = tid * 16 < n;                              // (partial tiles: the tail threads idle)
    const int p = p0 + tid * 16;                                 // pixel of the image (< 2^31)
    uint4 lab = make_uint4(0, 0, 0, 0);
    constexpr bool LAB = EXTRA == 2;
    if (LAB && ex.label_plane && mine) lab = __ldg(reinterpret_cast<const uint4*>(ex.label + b * hw + p));
    dt_mbar_wait(&full[s], (uint32_t)(k / kDtStages) & 1u);
    ExtrasRegs er;                                               // (EXTRA) footprint values, consumed after the hand-over
    bool own = false;
    int gidx = 0;
    if (mine) {
      uint32_t q[3][16];
#pragma unroll
      for (int c = 0; c < 3; ++c) quant16_staged(buf + (s * 3 + c) * kDtTile + tid * 16, q[c]);
      if (GRAY) {
        uint32_t w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
          w[j] = pack4<2>(luma_sum(q[0][4 * j], q[1][4 * j], q[2][4 * j]),
                          luma_sum(q[0][4 * j + 1], q[1][4 * j + 1], q[2][4 * j + 1]),
                          luma_sum(q[0][4 * j + 2], q[1][4 * j + 2], q[2][4 * j + 2]),
                          luma_sum(q[0][4 * j + 3], q[1][4 * j + 3], q[2][4 * j + 3]));
        __stcs(reinterpret_cast<uint4*>(gray + b * gray_batch_stride + p), make_uint4(w[0], w[1], w[2], w[3]));
      }
      if (RGB) {
        uint32_t w[12];
#pragma unroll
        for (int j = 0; j < 12; ++j)
          w[j] = pack4<0>(q[(4 * j) % 3][(4 * j) / 3], q[(4 * j + 1) % 3][(4 * j + 1) / 3],
                          q[(4 * j + 2) % 3][(4 * j + 2) / 3], q[(4 * j + 3) % 3][(4 * j + 3) / 3]);
        uint4* o = reinterpret_cast<uint4*>(rgb + (b * hw + p) * 3);
        __stcs(o, make_uint4(w[0], w[1], w[2], w[3]));
        __stcs(o + 1, make_uint4(w[4], w[5], w[6], w[7]));
        __stcs(o + 2, make_uint4(w[8], w[9], w[10], w[11]));
      }
      if (EXTRA) {
        gidx = p >> 4;
        if (LAB && ex.label_plane)
          __stcs(reinterpret_cast<uint4*>(ex.label_plane + b * ex.label_plane_stride + p), lab);
        const int y = ex.gpr_shift >= 0 ? (gidx >> ex.gpr_shift) : (gidx / ex.gpr);
        own = (y & 15) == 7;                                     // this thread owns a footprint: all of it is in the tile
        if (own) {
          // only the footprint's LOADS happen while the stage is held; the lerps and the scattered stores run after
          // the stage has been handed back, off the tile's critical path (the owners are 2 of the CTA's 8 warps)
          er.o = ((int64_t)(y >> 4)) * ex.gpr + (gidx - y * ex.gpr);
          if (ex.feat || ex.small_rgb) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const T* r0 = buf + (s * 3 + c) * kDtTile + tid * 16 + 7;
              er.px[c][0] = to_f32(r0[0]); er.px[c][1] = to_f32(r0[1]);
              er.px[c][2] = to_f32(r0[ex.W]); er.px[c][3] = to_f32(r0[ex.W + 1]);
            }
          }
          if (LAB && ex.label_small) {
            const uint8_t* r0 = ex.label + b * hw + p + 7;
            const uint8_t* r1 = r0 + ex.W;
            er.lab4 = (uint32_t)__ldg(r0) | ((uint32_t)__ldg(r0 + 1) << 8) | ((uint32_t)__ldg(r1) << 16) |
                      ((uint32_t)__ldg(r1 + 1) << 24);
          }
        }
      }
    }
    if (PW) {                                                    // this warp has read stage s: release it
      __syncwarp();
      if ((tid & 31) == 0) dt_mbar_arrive(&empty[s]);
    } else {
      __syncthreads();                                           // every thread has read stage s
      if (tid == 0) {
        // the refill is an async-proxy write over memory just read through the generic proxy
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        int nb = cb, ni = ci;                                    // (cb, ci) is tile k + 1 by now
        advance(nb, ni, (kDtStages - 1) * (int)gridDim.x);
        issue(k + kDtStages, nb, ni);
      }
    }
    if (EXTRA && own) {
      TailExtras fx = ex;
      fx.label_plane = nullptr;                                  // (stored above)
      if (!LAB) fx.label_small = nullptr;
      extras_finish<T>(fx, b, gidx, er);
    }
  }
}

// any H*W (no alignment assumptions): one pixel per thread
template <typename T>
__global__ void __launch_bounds__(256)
decode_tail_scalar_kernel(const T* __restrict__ img, uint8_t* __restrict__ rgb,
                          uint8_t* __restrict__ gray, int64_t hw, int64_t total,
                          int64_t gray_batch_stride, NormArgs na) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t b = i / hw, p = i - b * hw;
    const T* base = img + b * 3 * hw + p;
    const uint32_t r = Quant<T>::bits(to_f32(base[0]));
    const uint32_t g = Quant<T>::bits(to_f32(base[hw]));
    const uint32_t bl = Quant<T>::bits(to_f32(base[2 * hw]));
    if (gray) gray[b * gray_batch_stride + p] = (uint8_t)(luma_sum(r, g, bl) >> 16);
    if (rgb) {
      uint8_t* o = rgb + i * 3;
      o[0] = (uint8_t)r; o[1] = (uint8_t)g; o[2] = (uint8_t)bl;
    }
    if (na.out) {
      const uint32_t qv[3] = {r & 255u, g & 255u, bl & 255u};
#pragma unroll
      for (int c = 0; c < 3; ++c)
        na.out[(b * 3 + c) * hw + p] =
            __fdiv_rn(__fsub_rn(__fdiv_rn((float)qv[c], 255.f), na.mean[c]), na.stdv[c]);
    }
  }
}

// per-device attribute bookkeeping: cudaFuncSetAttribute applies to the CURRENT device only
template <typename F>
static void ensure_smem_attr(F kernel, int bytes, bool (&done)[kMaxDevices]) {
  const int dev = current_device();
  if (dev < 0 || dev >= kMaxDevices) { cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); return; }
  if (!done[dev]) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    done[dev] = true;
  }
}

template <typename T, bool RGB, bool GRAY, int EXTRA, int STAGES, int CTAS, bool PW = false>
static void launch_tma_one(const T* p, uint8_t* rgb, uint8_t* gray, int64_t hw, int tiles_per_img, int total,
                           int64_t gray_batch_stride, int tile_off, const TailExtras& ex, cudaStream_t st) {
  static bool attr[kMaxDevices] = {};
  const size_t smem = (size_t)STAGES * 3 * kDtTile * sizeof(T);
  auto k = decode_tail_tma_kernel<T, RGB, GRAY, EXTRA, STAGES, CTAS, PW>;
  ensure_smem_attr(k, (int)smem, attr);
  const int cap = CTAS * sm_count();
  const int grid = total < cap ? total : cap;
  k<<<grid, 256 + (PW ? 32 : 0), smem, st>>>(p, rgb, gray, hw, tiles_per_img, total, gray_batch_stride, tile_off, ex);
}

// pipeline shapes (stages x CTAs per SM) selectable through LDIFF_TUNE_DECODE_TAIL_TMA: smem in flight per SM is
// stages * CTAs * 3 planes * 4096 px * sizeof(T)
template <typename T> struct TmaShapes;
template <> struct TmaShapes<__nv_bfloat16> {
  template <bool RGB, bool GRAY, int EXTRA>
  static void launch(int variant, const __nv_bfloat16* p, uint8_t* rgb, uint8_t* gray, int64_t hw, int tpi, int total,
                     int64_t gbs, int toff, const TailExtras& ex, cudaStream_t st) {
    typedef __nv_bfloat16 T;
    switch (variant) {
      case 2: launch_tma_one<T, RGB, GRAY, EXTRA, 3, 3>(p, rgb, gray, hw, tpi, total, gbs, toff, ex, st); break;   // 216 KB
      case 3: launch_tma_one<T, RGB, GRAY, EXTRA, 2, 4>(p, rgb, gray, hw, tpi, total, gbs, toff, ex, st); break;   // 192 KB
      case 4: launch_tma_one<T, RGB, GRAY, EXTRA, 2, 3>(p, rgb, gray, hw, tpi, total, gbs, toff, ex, st); break;   // 144 KB
      case 5: launch_tma_one<T, RGB, GRAY, EXTRA, 3, 2>(p, rgb, gray, hw, tpi, total, gbs, toff, ex, st); break;   // 144 KB
      case 6: launch_tma_one<T, RGB, GRAY, EXTRA, 2, 2>(p, rgb, gray, hw, tpi, total, gbs, toff, ex, st); break;   //  96 KB
      case 7: launch_tma_one<T, RGB, GRAY, EXTRA, 2, 3, true>(p, rgb, gray, hw, tpi, total, gbs, toff, ex, st); break;   // producer warp
      case 8: launch_tma_one<T, RGB, GRAY, EXTRA, 3, 2, true>(p, rgb, gray, hw, tpi, total, gbs, toff, ex, st); break;
      case 9: launch_tma_one<T, RGB, GRAY, EXTRA, 2, 2, true>(p, rgb, gray, hw, tpi, total, gbs, toff, ex, st); break;
      case 10: launch_tma_one<T, RGB, GRAY, EXTRA, 4, 2, true>(p, rgb, gray, hw, tpi, total, gbs, toff, ex, st); break;
      case 11: launch_tma_one<T, RGB, GRAY, EXTRA, 2, 1>(p, rgb, gray, hw, tpi, total, gbs, toff, ex, st); break;  //  48 KB, one CTA per SM
      case 12: launch_tma_one<T, RGB, GRAY, EXTRA, 3, 1>(p, rgb, gray, hw, tpi, total, gbs, toff, ex, st); break;  //  72 KB
      case 13: launch_tma_one<T, RGB, GRAY, EXTRA, 4, 1>(p, rgb, gray, hw, tpi, total, gbs, toff, ex, st); break;  //  96 KB
      default: launch_tma_one<T, RGB, GRAY, EXTRA, 4, 2>(p, rgb, gray, hw, tpi, total, gbs, toff, ex, st); break;  // 192 KB
    }
  }
};
template <> struct TmaShapes<float> {
  template <bool RGB, bool GRAY, int EXTRA>
  static void launch(int variant, const float* p, uint8_t* rgb, uint8_t* gray, int64_t hw, int tpi, int total,
                     int64_t gbs, int toff, const TailExtras& ex, cudaStream_t st) {
    if (variant == 2 || variant == 3) launch_tma_one<float, RGB, GRAY, EXTRA, 2, 2>(p, rgb, gray, hw, tpi, total, gbs, toff, ex, st);
    else launch_tma_one<float, RGB, GRAY, EXTRA, 4, 1>(p, rgb, gray, hw, tpi, total, gbs, toff, ex, st);
  }
};

template <typename T>
static int launch_decode_tail(const void* img, uint8_t* rgb, uint8_t* gray, int B, int H, int W,
                              int64_t gray_batch_stride, NormArgs na, const TailExtras* exp, cudaStream_t st) {
  const int64_t hw = (int64_t)H * W;
  const int threads = 256;
  const bool vec_ok = (hw % 16 == 0) && aligned16(img) && aligned16(rgb) && aligned16(gray) &&
                      aligned16(na.out) && (gray_batch_stride % 16 == 0);
  const bool extra = exp != nullptr;
  if (extra && (!vec_ok || na.out || !gray)) return LDIFF_EUNSUPPORTED;      // (entry point checked the geometry)
  const TailExtras ex = extra ? *exp : TailExtras{};
  const int tma = tune_get(LDIFF_TUNE_DECODE_TAIL_TMA);
  if (vec_ok && !na.out && (hw % kDtTile) == 0 && tma > 0 && (int64_t)B * (hw / kDtTile) <= 0x7fffffff &&
      hw + kDtTile <= 0x7fffffff) {                      // (the kernel does its per-image pixel arithmetic in 32 bits)
    // extras: shift the tiles of an image by half a tile's rows (whole tiles of >= 16 rows need no shift) so that
    // rows 16e+7 and 16e+8 always share a tile; needs a tile to be a whole, power-of-two number of rows
    const int tile_rows = (W > 0 && kDtTile % W == 0) ? kDtTile / W : 0;
    int tile_off = 0;
    bool smem_extras = extra && tile_rows >= 2 && (tile_rows & (tile_rows - 1)) == 0;
    if (smem_extras) tile_off = ((tile_rows / 2) % 8) * W;
    if (extra && !smem_extras) goto reg_path;          // odd geometry: the register-staged kernel re-reads the footprints
    {
    const int tiles_per_img = (int)(hw / kDtTile) + (tile_off ? 1 : 0), total = B * tiles_per_img;
    const T* p = (const T*)img;
#define DTT(R, G, E) TmaShapes<T>::template launch<R, G, E>(tma, p, rgb, gray, hw, tiles_per_img, total, \
                                                            gray_batch_stride, tile_off, ex, st)
    if (extra && ex.label) { if (rgb) DTT(true, true, 2); else DTT(false, true, 2); }
    else if (extra) { if (rgb) DTT(true, true, 1); else DTT(false, true, 1); }
    else if (rgb && gray) DTT(true, true, 0);
    else if (gray) DTT(false, true, 0);
    else DTT(true, false, 0);
#undef DTT
    return check_launch();
    }
  }
reg_path:
  if (vec_ok) {
    if ((hw >> 4) > 0x7fffffff / 2 || B > 65535) return LDIFF_EUNSUPPORTED;
    const int gpi = (int)(hw >> 4);
    // blocks per image: cover it, but no more than ~8 resident blocks per SM over the batch
    // (balanced: every thread runs the same number of iterations, so the grid is one even wave)
    const int need = (gpi + threads - 1) / threads;
    const T* p = (const T*)img;
    // resident blocks per SM of THIS instantiation (register-limited: 4 for fp32, 6-7 for bf16)
#define DT(R, G, N, E)                                                                                \
  do {                                                                                                \
    static int occ = 0;                                                                               \
    if (occ == 0 && cudaOccupancyMaxActiveBlocksPerMultiprocessor(                                    \
                        &occ, decode_tail_vec16_kernel<T, R, G, N, E>, threads, 0) != cudaSuccess)    \
      occ = 4;                                                                                        \
    const int sms = (tune_get(LDIFF_TUNE_DECODE_TAIL_SMS) > 0 && tune_get(LDIFF_TUNE_DECODE_TAIL_SMS) < sm_count())   \
                        ? tune_get(LDIFF_TUNE_DECODE_TAIL_SMS) : sm_count();                          \
    int cap = (sms * (occ > 0 ? occ : 4)) / B;                                                        \
    if (cap < 1) cap = 1;                                                                             \
    const int iters = (need + cap - 1) / cap;                                                         \
    const dim3 grid((need + iters - 1) / iters, B);                                                   \
    decode_tail_vec16_kernel<T, R, G, N, E><<<grid, threads, 0, st>>>(p, rgb, gray, hw, gpi,          \
                                                                      gray_batch_stride, na, ex);     \
  } while (0)
    if (extra) {
      if (rgb) DT(true, true, false, true);
      else DT(false, true, false, true);
    } else if (na.out) {
      if (rgb && gray) DT(true, true, true, false);
      else if (gray) DT(false, true, true, false);
      else if (rgb) DT(true, false, true, false);
      else DT(false, false, true, false);
    } else {
      if (rgb && gray) DT(true, true, false, false);
      else if (gray) DT(false, true, false, false);
      else DT(true, false, false, false);
    }
#undef DT
  } else {
    const int64_t total = hw * B;
    decode_tail_scalar_kernel<T><<<grid_for(total, threads, 8), threads, 0, st>>>(
        (const T*)img, rgb, gray, hw, total, gray_batch_stride, na);
  }
  return check_launch();
}

}  // namespace ldiff

using namespace ldiff;

static int decode_tail_entry(const void* img, uint8_t* rgb_hwc, uint8_t* gray, float* model_input,
                             const float* mean3, const float* std3, int B, int H, int W,
                             int64_t gray_batch_stride, int dtype, const TailExtras* ex, void* stream) {
  if (!img || (!rgb_hwc && !gray && !model_input) || B < 0 || H < 0 || W < 0) return LDIFF_EINVAL;
  if (gray && gray_batch_stride < (int64_t)H * W) return LDIFF_EINVAL;
  if (model_input && (!mean3 || !std3)) return LDIFF_EINVAL;
  if (B == 0 || H == 0 || W == 0) return LDIFF_OK;
  NormArgs na;
  na.out = model_input;
  for (int c = 0; c < 3; ++c) {
    na.mean[c] = model_input ? mean3[c] : 0.f;          // host pointers: three floats each
    na.stdv[c] = model_input ? std3[c] : 1.f;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == LDIFF_F32)
    return launch_decode_tail<float>(img, rgb_hwc, gray, B, H, W, gray_batch_stride, na, ex, st);
  if (dtype == LDIFF_BF16)
    return launch_decode_tail<__nv_bfloat16>(img, rgb_hwc, gray, B, H, W, gray_batch_stride, na, ex, st);
  return LDIFF_EUNSUPPORTED;
}

extern "C" int ldiff_decode_tail_gray(const void* img, uint8_t* rgb_hwc, uint8_t* gray, int B, int H,
                                      int W, int64_t gray_batch_stride, int dtype, void* stream) {
  return decode_tail_entry(img, rgb_hwc, gray, nullptr, nullptr, nullptr, B, H, W, gray_batch_stride,
                           dtype, nullptr, stream);
}

extern "C" int ldiff_decode_tail_model_input(const void* img, uint8_t* rgb_hwc, uint8_t* gray,
                                             float* model_input, const float* host_mean3,
                                             const float* host_std3, int B, int H, int W,
                                             int64_t gray_batch_stride, int dtype, void* stream) {
  if (!model_input) return LDIFF_EINVAL;
  return decode_tail_entry(img, rgb_hwc, gray, model_input, host_mean3, host_std3, B, H, W,
                           gray_batch_stride, dtype, nullptr, stream);
}

extern "C" int ldiff_decode_tail_fused(const void* img, uint8_t* rgb_hwc, uint8_t* gray, int B, int H, int W,
                                       int64_t gray_batch_stride, int dtype, void* feat, int feat_dtype,
                                       int feat_ctot, int feat_channel, void* small_rgb, const uint8_t* label,
                                       uint8_t* label_plane, int64_t label_plane_stride, uint8_t* label_small,
                                       void* stream) {
  if (!img || !gray || B < 0 || H < 0 || W < 0) return LDIFF_EINVAL;
  if (!feat && !small_rgb && !label_plane && !label_small) return LDIFF_EINVAL;       // use ldiff_decode_tail_gray
  if ((label_plane || label_small) && !label) return LDIFF_EINVAL;
  if (feat && (feat_channel < 0 || feat_channel >= feat_ctot)) return LDIFF_EINVAL;
  if (feat && feat_dtype != dtype && feat_dtype != LDIFF_F32) return LDIFF_EUNSUPPORTED;
  if (label_plane && label_plane_stride < (int64_t)H * W) return LDIFF_EINVAL;
  if (B == 0 || H == 0 || W == 0) return LDIFF_OK;
  if ((H % 16) || (W % 16)) return LDIFF_EUNSUPPORTED;  // the 16x down-sample is the 2x2 footprint only then
  if (!aligned16(label) || !aligned16(label_plane) || (label_plane && (label_plane_stride % 16))) return LDIFF_EALIGN;
  TailExtras ex;
  ex.feat = feat; ex.feat_ctot = feat_ctot; ex.feat_ch = feat_channel;
  ex.feat_f32 = (feat && feat_dtype == LDIFF_F32 && dtype != LDIFF_F32) ? 1 : 0;
  ex.small_rgb = small_rgb; ex.label = label; ex.label_plane = label_plane;
  ex.label_plane_stride = label_plane_stride; ex.label_small = label_small;
  ex.W = W; ex.fh = H / 16; ex.gpr = W / 16; ex.gpr_shift = -1;
  for (int sft = 0; sft < 31; ++sft)
    if ((1 << sft) == ex.gpr) ex.gpr_shift = sft;
  return decode_tail_entry(img, rgb_hwc, gray, nullptr, nullptr, nullptr, B, H, W, gray_batch_stride, dtype, &ex,
                           stream);
}
