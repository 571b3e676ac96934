// a-5 classifier-head contraction on the 5th-gen tensor cores (sm_100a):
//   logits[b, k, p] = bias[k] + sum_c W[k, c] * feat[b, c, p]      (conductor.py:127)
// bf16 operands, fp32 accumulation in TMEM.
//
// GEMM view per CTA: D[128 pixels, 16 classes] = A[128, Cin] * B[16, Cin]^T with
// one tcgen05.mma (M=128, N=16, K=16) per 16 channels, issued by a single thread.
//
//  * The feature map is NCHW: pixels are contiguous, channels strided, i.e. the
//    A operand arrives "M-major".  Instead of an MN-major descriptor the tile is
//    transposed on its way into shared memory: each thread gathers 8 consecutive
//    channels of one pixel (coalesced 2-byte loads across the warp) and writes
//    one 16-byte chunk of the canonical K-major SWIZZLE_128B layout (row = pixel,
//    128 bytes = 64 channels per row, chunk index XOR (row & 7)).  W is K-major
//    already and is copied with 16-byte loads into the same layout.
//  * The whole K extent fits in shared memory (Cin/64 blocks of 16 KB + 2 KB), so
//    there is no pipeline: fill, fence.proxy.async, one elected thread issues
//    Cin/16 MMAs and a tcgen05.commit onto an mbarrier, four warps read the
//    accumulator back with tcgen05.ld (lane = pixel, 16 columns = classes), add
//    the bias and store coalesced fp32 logits.
//
// The problem is tiny (8 x 1024 pixels x 256 x 11: 46 MFLOP, 4 MB): the tensor
// core is there to take the math off the critical path, the kernel is bound by
// the latency of one pass over its 64 KB tile.
#include <cuda.h>          // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)

#include <cstdlib>

#include "head_common.cuh"

namespace ldiff {

namespace tc {

constexpr int kTileM = 128;     // pixels per CTA  (UMMA M)
constexpr int kTileN = 16;      // padded classes  (UMMA N)
constexpr int kBlockK = 64;     // channels per 128-byte swizzle row
constexpr int kUmmaK = 16;      // channels per tcgen05.mma (bf16)
constexpr int kThreads = 256;
constexpr uint32_t kTmemCols = 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (sm_100 format):
//   [0,14) start address >> 4, [16,30) LBO >> 4 (unused for swizzled K-major),
//   [32,46) SBO >> 4 = 1024 B between 8-row groups, [46,48) version = 1,
//   [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// instruction descriptor, kind::f16: D=f32 (bit 4), A=B=bf16 (bits 7, 10), both
// K-major (bits 15, 16 = 0), N>>3 at [17,23), M>>4 at [24,29)
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kTileN >> 3) << 17) |
                            ((uint32_t)(kTileM >> 4) << 24);

__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(kIdesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "TC_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra TC_DONE;\n\t"
      "bra TC_WAIT;\n\t"
      "TC_DONE:\n\t}" ::"r"(bar), "r"(parity)
      : "memory");
}

// grid = (hw / 128, B); dynamic smem = 1024 (alignment slack) + nkb * (16 KB + 2 KB)
__global__ void __launch_bounds__(kThreads, 1)
head_logits_tc_kernel(const __nv_bfloat16* __restrict__ feat, const __nv_bfloat16* __restrict__ weight,
                      const float* __restrict__ bias, float* __restrict__ logits, int Cin, int K, int hw,
                      unsigned long long* __restrict__ clear, int n_clear) {
  extern __shared__ uint8_t smem_raw[];
  clear_counters(clear, n_clear);
  __shared__ __align__(8) uint64_t mma_bar;
  __shared__ uint32_t tmem_base_slot;

  // SWIZZLE_128B atoms must be 1024-byte aligned
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int nkb = Cin / kBlockK;
  uint8_t* sA = smem;                                   // [nkb][128 rows][128 B]
  uint8_t* sB = smem + (size_t)nkb * kTileM * 128;      // [nkb][ 16 rows][128 B]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int p0 = blockIdx.x * kTileM;

  if (warp == 0) {                                      // TMEM allocation (one warp, .sync.aligned)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_base_slot)), "n"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 32) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mma_bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  // ---- A: transpose [c][p] -> K-major rows.  item = (row r, 16-byte chunk j of the full K extent)
  const __nv_bfloat16* fb = feat + (int64_t)b * Cin * hw + p0;
  const int chunks_per_row = Cin / 8;
  for (int item = threadIdx.x; item < kTileM * chunks_per_row; item += kThreads) {
    const int r = item % kTileM;                         // consecutive threads -> consecutive pixels
    const int j = item / kTileM;                         // channels 8j .. 8j+7
    const __nv_bfloat16* src = fb + (int64_t)(8 * j) * hw + r;
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint16_t lo = __ldg(reinterpret_cast<const uint16_t*>(src + (int64_t)(2 * q) * hw));
      const uint16_t hi = __ldg(reinterpret_cast<const uint16_t*>(src + (int64_t)(2 * q + 1) * hw));
      w[q] = (uint32_t)lo | ((uint32_t)hi << 16);
    }
    const int kb = j >> 3, ch = j & 7;
    uint8_t* dst = sA + ((size_t)kb * kTileM + r) * 128 + ((ch ^ (r & 7)) << 4);
    *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
  }
  // ---- B: weights [K][Cin] (K-major), rows >= K are zero
  for (int item = threadIdx.x; item < kTileN * chunks_per_row; item += kThreads) {
    const int n = item / chunks_per_row, j = item - n * chunks_per_row;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (n < K) v = __ldg(reinterpret_cast<const uint4*>(weight + (int64_t)n * Cin) + j);
    const int kb = j >> 3, ch = j & 7;
    *reinterpret_cast<uint4*>(sB + ((size_t)kb * kTileN + n) * 128 + ((ch ^ (n & 7)) << 4)) = v;
  }
  // generic-proxy writes -> visible to the async proxy (tensor core) before the MMAs
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_slot;

  // ---- MMA issue: one thread
  if (threadIdx.x == 0) {
    uint32_t acc = 0;
    for (int kb = 0; kb < nkb; ++kb) {
      const uint32_t a0 = smem_u32(sA + (size_t)kb * kTileM * 128);
      const uint32_t b0 = smem_u32(sB + (size_t)kb * kTileN * 128);
#pragma unroll
      for (int k = 0; k < kBlockK / kUmmaK; ++k) {        // advance 32 bytes inside the swizzle atom
        mma_bf16(tmem_d, make_desc(a0 + k * kUmmaK * 2), make_desc(b0 + k * kUmmaK * 2), acc);
        acc = 1;
      }
    }
    // arrives on the mbarrier when every MMA above has completed
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(&mma_bar)) : "memory");
  }

  // ---- epilogue: warps 0..3, warp w owns TMEM lanes 32w..32w+31 (= pixels)
  if (warp < 4) {
    mbar_wait(smem_u32(&mma_bar), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v[16];
    const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const int p = p0 + warp * 32 + lane;
    float* out = logits + (int64_t)b * K * hw + p;
#pragma unroll
    for (int k = 0; k < kTileN; ++k)
      if (k < K) out[(int64_t)k * hw] = __uint_as_float(v[k]) + (bias ? __ldg(bias + k) : 0.f);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(kTmemCols)
                 : "memory");
  }
}


// ----------------------------------------------------------------------------
// The same contraction with a Blackwell-native load side: tensor-map TMA (cp.async.bulk.tensor -> SASS UTMALDG)
// brings the NCHW feature tile in AS IT LIES IN MEMORY — pixels contiguous, i.e. the A operand is MN-major — and the
// tensor core reads it through an MN-major SWIZZLE_128B descriptor, so nothing is transposed by threads:
//   A, per 64-channel block kb: two boxes {64 pixels, 64 channels} of 8 KB; a box row is one channel's 64 pixels
//      (128 B), TMA's 128-byte swizzle XORs the 16-byte chunk index with (row & 7) — exactly UMMA's canonical
//      MN-major SW128 atom (64 MN x 8 K); LBO = 8 KB (next 64 pixels), SBO = 1 KB (next 8 channels)
//   B, per kb: one box {64 channels, 16 classes} of 2 KB from W[K, Cin] (K-major as before; rows >= K are
//      out of bounds and arrive as zeros)
// One mbarrier per kb: the issuing thread starts the MMAs of a block as soon as ITS 18 KB have landed, while the
// later blocks are still in flight (the round-1 kernel filled the whole 72 KB with 2-byte loads before the
// first MMA).  Epilogue unchanged.
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);
  d |= (uint64_t)(8192u >> 4) << 16;                  // LBO: the next 64-pixel atom
  d |= (uint64_t)(1024u >> 4) << 32;                  // SBO: the next group of 8 channels
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;                             // SWIZZLE_128B
  return d;
}
constexpr uint32_t kIdescMN = kIdesc | (1u << 15);    // A is MN-major

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void mma_bf16_mn(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(kIdescMN), "r"(accumulate)
      : "memory");
}

constexpr int kMaxKb = 8;                              // Cin <= 512

// mbarrier wait that cannot hang the GPU: a tensor map the hardware rejects would otherwise leave the block
// spinning forever; after ~0.3 s the kernel traps (the launch then fails loudly at the next synchronisation)
__device__ __forceinline__ void mbar_wait_or_trap(uint32_t bar, uint32_t parity) {
  for (uint32_t i = 0; i < (1u << 22); ++i) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return;
    __nanosleep(64);
  }
  asm volatile("trap;");
}

// grid = (hw / 128, B), 128 threads; dynamic smem = 1024 + nkb * (16 KB + 2 KB)
__global__ void __launch_bounds__(128, 1)
head_logits_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const float* __restrict__ bias, float* __restrict__ logits, int Cin, int K, int hw,
                       unsigned long long* __restrict__ clear, int n_clear) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full[kMaxKb];
  __shared__ __align__(8) uint64_t mma_bar;
  __shared__ uint32_t tmem_base_slot;
  clear_counters(clear, n_clear);
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int nkb = Cin / kBlockK;
  uint8_t* sA = smem;                                   // [nkb][2 atoms][64 channels][128 B]
  uint8_t* sB = smem + (size_t)nkb * 16384;             // [nkb][16 classes][128 B]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y, p0 = blockIdx.x * kTileM;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_base_slot)), "n"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 32) {
    for (int kb = 0; kb < nkb; ++kb)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[kb])) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mma_bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_slot;

  if (threadIdx.x == 0) {                               // ---- producer + MMA issuer: one thread
    for (int kb = 0; kb < nkb; ++kb) {
      const uint32_t bar = smem_u32(&full[kb]);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(16384u + 2048u) : "memory");
      tma_load_3d(sA + (size_t)kb * 16384, &tmA, p0, kb * kBlockK, b, bar);
      tma_load_3d(sA + (size_t)kb * 16384 + 8192, &tmA, p0 + 64, kb * kBlockK, b, bar);
      tma_load_2d(sB + (size_t)kb * 2048, &tmB, kb * kBlockK, 0, bar);
    }
    uint32_t acc = 0;
    for (int kb = 0; kb < nkb; ++kb) {
      mbar_wait_or_trap(smem_u32(&full[kb]), 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a0 = smem_u32(sA + (size_t)kb * 16384);
      const uint32_t b0 = smem_u32(sB + (size_t)kb * 2048);
#pragma unroll
      for (int k = 0; k < kBlockK / kUmmaK; ++k) {        // A: 16 channels = two 8-channel groups of 1 KB; B: 32 bytes
        mma_bf16_mn(tmem_d, make_desc_mn(a0 + k * 2048), make_desc(b0 + k * kUmmaK * 2), acc);
        acc = 1;
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(&mma_bar)) : "memory");
  }
  __syncwarp();

  // ---- epilogue: warp w owns TMEM lanes 32w..32w+31 (= pixels)
  mbar_wait_or_trap(smem_u32(&mma_bar), 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t v[16];
  const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  const int p = p0 + warp * 32 + lane;
  float* out = logits + (int64_t)b * K * hw + p;
#pragma unroll
  for (int k = 0; k < kTileN; ++k)
    if (k < K) out[(int64_t)k * hw] = __uint_as_float(v[k]) + (bias ? __ldg(bias + k) : 0.f);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(kTmemCols)
                 : "memory");
  }
}

// ----------------------------------------------------------------------------
// Cell form on the same tensor-core path (conductor.py:218-221 for a whole batch): the instance features are
// [n_total, Cin] row-major, i.e. the A operand is K-major as it lies in memory — one box {64 channels, 128 instances}
// per 64-channel block lands in the canonical K-major SWIZZLE_128B layout.  D[128 instances, 16 classes]; the thread
// that reads instance i's 16 accumulator columns adds the bias, takes the pinned decision (cell_decide) and writes
// the instance's LUT byte.  Replaces a CUDA-core kernel that spent 5.5 M warp instructions (as many as a decode
// tail) on 18 MFLOP.
// grid = ceil(n_total / 128), 128 threads; dynamic smem = 1024 + nkb * (16 KB + 2 KB)
__global__ void __launch_bounds__(128, 1)
cell_classify_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const float* __restrict__ bias, const int32_t* __restrict__ ids, uint8_t* __restrict__ lut,
                         int lut_size, int64_t lut_stride, float* __restrict__ logits_out, int n_per_image,
                         int n_total, int Cin, int K, int* __restrict__ status,
                         unsigned long long* __restrict__ clear, int n_clear) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full[kMaxKb];
  __shared__ __align__(8) uint64_t mma_bar;
  __shared__ uint32_t tmem_base_slot;
  clear_counters(clear, n_clear);
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int nkb = Cin / kBlockK;
  uint8_t* sA = smem;                                   // [nkb][128 instances][128 B]
  uint8_t* sB = smem + (size_t)nkb * 16384;             // [nkb][16 classes][128 B]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * kTileM;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_base_slot)), "n"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 32) {
    for (int kb = 0; kb < nkb; ++kb)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[kb])) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mma_bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_slot;

  if (threadIdx.x == 0) {                               // ---- producer + MMA issuer: one thread
    for (int kb = 0; kb < nkb; ++kb) {
      const uint32_t bar = smem_u32(&full[kb]);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(16384u + 2048u) : "memory");
      tma_load_2d(sA + (size_t)kb * 16384, &tmA, kb * kBlockK, row0, bar);     // rows >= n_total arrive as zeros
      tma_load_2d(sB + (size_t)kb * 2048, &tmB, kb * kBlockK, 0, bar);
    }
    uint32_t acc = 0;
    for (int kb = 0; kb < nkb; ++kb) {
      mbar_wait_or_trap(smem_u32(&full[kb]), 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a0 = smem_u32(sA + (size_t)kb * 16384);
      const uint32_t b0 = smem_u32(sB + (size_t)kb * 2048);
#pragma unroll
      for (int k = 0; k < kBlockK / kUmmaK; ++k) {        // advance 32 bytes inside the swizzle atom
        mma_bf16(tmem_d, make_desc(a0 + k * kUmmaK * 2), make_desc(b0 + k * kUmmaK * 2), acc);
        acc = 1;
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(&mma_bar)) : "memory");
  }
  __syncwarp();

  // ---- epilogue: warp w owns TMEM lanes 32w..32w+31 (= instances)
  mbar_wait_or_trap(smem_u32(&mma_bar), 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t a[16];
  const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]),
        "=r"(a[8]), "=r"(a[9]), "=r"(a[10]), "=r"(a[11]), "=r"(a[12]), "=r"(a[13]), "=r"(a[14]), "=r"(a[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  const int inst = row0 + warp * 32 + lane;
  if (inst < n_total) {
    float v[kTileN];
#pragma unroll
    for (int k = 0; k < kTileN; ++k) v[k] = (k < K) ? __uint_as_float(a[k]) + (bias ? __ldg(bias + k) : 0.f) : -INFINITY;
    if (logits_out)
      for (int k = 0; k < K; ++k) logits_out[(int64_t)inst * K + k] = v[k];
    const int cls = cell_decide<kTileN>(v, K);
    const int img = inst / n_per_image;
    const int id = __ldg(ids + (inst - img * n_per_image));
    if (id >= 0 && id < lut_size) lut[(int64_t)img * lut_stride + id] = (uint8_t)cls;
    else atomicOr(status, LDIFF_STATUS_INST_RANGE);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(kTmemCols)
                 : "memory");
  }
}

}  // namespace tc

// cuTensorMapEncodeTiled without linking libcuda: fetched once through the runtime
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    cudaGetLastError();
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// tensor-map TMA form; LDIFF_EUNSUPPORTED when the shape / alignment / driver cannot take it
static int launch_head_logits_tma(const void* feat, const void* weight, const float* bias, float* logits, int B,
                                  int Cin, int K, int hw, int64_t* clear, int n_clear, cudaStream_t st) {
  using namespace tc;
  static const bool off = [] { const char* e = getenv("LDIFF_HEAD_TMA"); return e && e[0] == '0'; }();
  if (off || Cin / kBlockK > kMaxKb || !aligned16(feat) || !aligned16(weight) || (hw % 8) || (Cin % 8))
    return LDIFF_EUNSUPPORTED;
  EncodeTiledFn enc = encode_tiled();
  if (!enc) return LDIFF_EUNSUPPORTED;
  alignas(64) CUtensorMap tmA, tmB;
  const cuuint32_t ones[3] = {1, 1, 1};
  {
    const cuuint64_t dims[3] = {(cuuint64_t)hw, (cuuint64_t)Cin, (cuuint64_t)B};
    const cuuint64_t strides[2] = {(cuuint64_t)hw * 2, (cuuint64_t)Cin * hw * 2};
    const cuuint32_t box[3] = {64, 64, 1};
    if (enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(feat), dims, strides, box, ones,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return LDIFF_EUNSUPPORTED;
  }
  {
    const cuuint64_t dims[2] = {(cuuint64_t)Cin, (cuuint64_t)K};
    const cuuint64_t strides[1] = {(cuuint64_t)Cin * 2};
    const cuuint32_t box[2] = {64, (cuuint32_t)kTileN};            // classes K .. 15 are out of bounds: zero-filled
    if (enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(weight), dims, strides, box, ones,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return LDIFF_EUNSUPPORTED;
  }
  const int nkb = Cin / kBlockK;
  const size_t smem = 1024 + (size_t)nkb * (16384 + 2048);
  static bool attr_set[kMaxDevices] = {};
  const int dev = current_device();
  if (dev < 0 || dev >= kMaxDevices || !attr_set[dev]) {
    cudaFuncSetAttribute(head_logits_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (dev >= 0 && dev < kMaxDevices) attr_set[dev] = true;
  }
  head_logits_tma_kernel<<<dim3(hw / kTileM, B), 128, smem, st>>>(tmA, tmB, bias, logits, Cin, K, hw,
                                                                  reinterpret_cast<unsigned long long*>(clear), n_clear);
  return check_launch();
}

// tensor-core cell classifier; LDIFF_EUNSUPPORTED when the shape / alignment / driver cannot take it (the caller
// then runs the CUDA-core kernel).  LDIFF_CELL_TC=0 forces the fall-back.
int launch_cell_classify_tc(const void* feats, const void* weight, const float* bias, const int32_t* ids, uint8_t* lut,
                            int lut_size, int64_t lut_stride, float* logits_out, int n_per_image, int n_total, int Cin,
                            int K, int64_t* clear, int n_clear, int* status, cudaStream_t st) {
  using namespace tc;
  static const bool off = [] { const char* e = getenv("LDIFF_CELL_TC"); return e && e[0] == '0'; }();
  if (off || K > kTileN || (Cin % kBlockK) != 0 || Cin / kBlockK > kMaxKb || !aligned16(feats) || !aligned16(weight) ||
      n_total < 1)
    return LDIFF_EUNSUPPORTED;
  EncodeTiledFn enc = encode_tiled();
  if (!enc) return LDIFF_EUNSUPPORTED;
  alignas(64) CUtensorMap tmA, tmB;
  const cuuint32_t ones[2] = {1, 1};
  const cuuint64_t strides[1] = {(cuuint64_t)Cin * 2};
  {
    const cuuint64_t dims[2] = {(cuuint64_t)Cin, (cuuint64_t)n_total};
    const cuuint32_t box[2] = {64, (cuuint32_t)kTileM};            // instances past n_total: zero-filled
    if (enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(feats), dims, strides, box, ones,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return LDIFF_EUNSUPPORTED;
  }
  {
    const cuuint64_t dims[2] = {(cuuint64_t)Cin, (cuuint64_t)K};
    const cuuint32_t box[2] = {64, (cuuint32_t)kTileN};
    if (enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(weight), dims, strides, box, ones,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return LDIFF_EUNSUPPORTED;
  }
  const int nkb = Cin / kBlockK;
  const size_t smem = 1024 + (size_t)nkb * (16384 + 2048);
  static bool attr_set[kMaxDevices] = {};
  const int dev = current_device();
  if (dev < 0 || dev >= kMaxDevices || !attr_set[dev]) {
    cudaFuncSetAttribute(cell_classify_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (dev >= 0 && dev < kMaxDevices) attr_set[dev] = true;
  }
  cell_classify_tma_kernel<<<(n_total + kTileM - 1) / kTileM, 128, smem, st>>>(
      tmA, tmB, bias, ids, lut, lut_size, lut_stride, logits_out, n_per_image, n_total, Cin, K, status,
      reinterpret_cast<unsigned long long*>(clear), n_clear);
  return check_launch();
}

int launch_head_logits_tc(const void* feat, const void* weight, const float* bias, float* logits,
                          int B, int Cin, int K, int hw, int64_t* clear, int n_clear, cudaStream_t st) {
  using namespace tc;
  if (K > kTileN || (hw % kTileM) != 0 || (Cin % kBlockK) != 0 || B > 65535) return LDIFF_EUNSUPPORTED;
  {
    const int rc = launch_head_logits_tma(feat, weight, bias, logits, B, Cin, K, hw, clear, n_clear, st);
    if (rc != LDIFF_EUNSUPPORTED) return rc;
  }
  if (!aligned16(weight) || (reinterpret_cast<uintptr_t>(feat) & 1)) return LDIFF_EUNSUPPORTED;
  const int nkb = Cin / kBlockK;
  const size_t smem = 1024 + (size_t)nkb * (kTileM + kTileN) * 128;
  if (smem > 200 * 1024) return LDIFF_EUNSUPPORTED;
  static bool attr_set[kMaxDevices] = {};              // the attribute belongs to the CURRENT device
  const int dev = current_device();
  if (dev < 0 || dev >= kMaxDevices || !attr_set[dev]) {
    cudaFuncSetAttribute(head_logits_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (dev >= 0 && dev < kMaxDevices) attr_set[dev] = true;
  }
  dim3 grid(hw / kTileM, B);
  head_logits_tc_kernel<<<grid, kThreads, smem, st>>>((const __nv_bfloat16*)feat,
                                                       (const __nv_bfloat16*)weight, bias, logits, Cin, K, hw,
                                                       reinterpret_cast<unsigned long long*>(clear), n_clear);
  return check_launch();
}

}  // namespace ldiff
