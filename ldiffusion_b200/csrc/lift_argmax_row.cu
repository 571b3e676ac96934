// a-5 lift + argmax, ROW form of the envelope kernel (sm_100a): conductor.py:135 + segmentor.py:536 for the
// x32 horizontal lift of the path (1024 <- 32, 512 <- 16).
#include "head_common.cuh"
#include "hist.cuh"

namespace ldiff {

// ----------------------------------------------------------------------------
// The envelope argument of lift_argmax_env.cu turned by 90 degrees.  For a fixed output ROW y the K lifted
// logits are, inside one source cell, K LINES in the horizontal weight l:
//     v_k(l) = A_k(c) + l * (A_k(c+1) - A_k(c)),     A_k(c) = hy0 * p_k[i0y][c] + hy1 * p_k[i1y][c]
// (bilinear interpolation is linear in l whichever axis is lifted first), so a thread that owns one output row
// sweeps the upper envelope along x instead of down a column.  What that buys:
//  * the 32 pixels of a cell are 32 CONSECUTIVE BYTES of the mask: a run (first column, class) is XORed into eight
//    word registers with byte masks (4 instructions per word) and the cell leaves as two 16-byte stores — the
//    column form replays its segments row by row, 17 instructions and one 2-byte store per row and column pair;
//  * the vertical lift that sets the lines up reads its four source values as two 8-byte shared-memory loads at
//    compile-time offsets (packed FMUL2 / FFMA2 give T and U together): ~90 instructions per cell and row against
//    ~350 per column and band;
//  * with the x32 lift the weight of column j is exactly (j + 0.5) / 32 in fp32, so "the first column behind a
//    run" is arithmetic (one FFMA + ceil), not a search in a weight table; no limit on the segments per cell.
// A "chunk" is the part of a row that shares one source-column pair: x in [32c - 16, 32c + 16), c = 0 .. w (the two
// half chunks at the image border are flat lines: U = T).  A thread sweeps the two chunks 2p and 2p + 1 of its row
// (64 bytes), a block owns 256 rows x one chunk pair and stages the <= 12 source rows x 3 source columns of all K
// classes it needs in shared memory.
//
// Why a run is safe (M = the largest |A_k| of the chunk, E_k(l) the exact bilinear value through the fp32 sources
// at the row's fp32 weights):
//  * the reference's value (horizontal lerp, then vertical: oracle/bilinear.py::lift_spec) is within 4 * 2^-24 * M
//    of E_k; the swept v' = fma(l, fl(U - T), T) with T, U rounded vertical lerps is within 6 * 2^-24 * M, so a lead
//    seen here is within 20 * 2^-24 * M of the reference's lead at that pixel;
//  * extrapolating a lead over l' - l <= 1 with rounded slope differences adds <= 8 * 2^-24 * M, the approximate
//    reciprocals a relative 2^-21 on l' - l, i.e. <= 2^-19 * M on the lead (the run is also shortened by 2^-12);
//  so with gap = 1e-5 + 2^-16 * M (four times the sum above) every pixel of a run has a reference top-2 gap above 1e-5
//  and the same winner, where argmax(softmax) == argmax (oracle/head.py).  Every other pixel goes through the pinned
//  softmax on the reference's own roundings, exactly as in the column form: the fast path only decides WHO is safe.
constexpr int kRowThreads = 256;   // output rows per block, one per thread
constexpr int kRowSrc = 12;        // source rows 256 output rows may span (x32 vertical lift: 10)
constexpr int kRowF = 32;          // the horizontal lift factor this form is written for

__device__ __forceinline__ float rcp_approx_row(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// 0xffffffff << n with PTX semantics: n >= 32 gives 0
__device__ __forceinline__ uint32_t ones_from_bit(int n) {
  uint32_t r;
  asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(0xffffffffu), "r"(n));
  return r;
}

// HIST: the a-6 confusion histogram of the mask against gt in the same pass — the eight class words of a chunk meet
// the chunk's 32 ground-truth bytes (two 16-byte loads) in registers; uncertain pixels are entered as class 0 and
// moved to their resolved class by the epilogue.  PUSH: with the NVLink peer push as the kernel's tail (hist.cuh).
template <int K, bool HIST, bool PUSH>
__global__ void __launch_bounds__(kRowThreads, 1024 / kRowThreads)   // 64 registers
lift_argmax_row_kernel(const float* __restrict__ logits, uint8_t* __restrict__ mask, const uint8_t* __restrict__ gt,
                       unsigned long long* __restrict__ C, AxisH ay, AxisH ax, int npairs, int nrb, int B,
                       int* __restrict__ status, XchgPush px) {
  __shared__ __align__(16) float s_src[K * kRowSrc * 4];   // [class][source row][c0, c1, c1, c2]
  __shared__ uint32_t s_q[kRowThreads / 32][kQueue];
  __shared__ int s_qn[kRowThreads / 32];
  __shared__ uint32_t s_hist[HIST ? (K + 1) * K * 32 : 1];
  __shared__ unsigned long long s_step;
  BlockHist<32, true> h;
  if (HIST) h.init(s_hist, K);                           // (the first item's barriers order it before any update)
  if (PUSH && threadIdx.x == 0) s_step = xchg_step_of_launch(px);
  const int warp_in_block = threadIdx.x >> 5;
  const int plane = ay.in * ax.in;
  const int items = npairs * nrb * B;
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int p = item % npairs, rest = item / npairs;
    const int rb = rest % nrb, b = rest / nrb;
    __syncthreads();                                     // the previous item's sweeps and epilogue are done with shared memory
    const int Y0 = rb * kRowThreads;
    const int sy0 = tap(ay, Y0).i0;
    const int nsr = min(tap(ay, min(Y0 + kRowThreads, ay.out) - 1).i1 - sy0 + 1, kRowSrc);   // (host: it fits)
    const int c0 = max(2 * p - 1, 0), c1 = min(2 * p, ax.in - 1), c2 = min(2 * p + 1, ax.in - 1);
    const float* lb = logits + (int64_t)b * K * plane;
    for (int i = threadIdx.x; i < K * nsr * 4; i += kRowThreads) {
      const int q = i & 3, kr = i >> 2;
      const int k = kr / nsr, r = kr - k * nsr;
      const int c = q == 0 ? c0 : (q == 3 ? c2 : c1);
      s_src[(k * kRowSrc + r) * 4 + q] = __ldg(lb + k * plane + (sy0 + r) * ax.in + c);
    }
    if (threadIdx.x < kRowThreads / 32) s_qn[threadIdx.x] = 0;
    __syncthreads();

    const int y = Y0 + (int)threadIdx.x;
    const bool active = y < ay.out;
    const TapH ty = tap(ay, min(y, ay.out - 1));
    const int ro0 = (ty.i0 - sy0) * 4, ro1 = (ty.i1 - sy0) * 4;
    const float2 hy0 = make_float2(ty.l0, ty.l0), hy1 = make_float2(ty.l1, ty.l1);
    uint8_t* orow = mask + ((int64_t)b * ay.out + min(y, ay.out - 1)) * ax.out;
    const uint8_t* grow = HIST ? gt + ((int64_t)b * ay.out + min(y, ay.out - 1)) * ax.out : nullptr;
    bool ovf = false;
    auto enqueue = [&](int x) -> bool {                  // false: the warp's queue is full
      const int slot = atomicAdd(&s_qn[warp_in_block], 1);
      if (slot >= kQueue) return false;
      s_q[warp_in_block][slot] = (threadIdx.x << 16) | (uint32_t)x;
      return true;
    };

#pragma unroll 1
    for (int cc = 0; cc < 2; ++cc) {
      const int c = 2 * p + cc;
      if (c > ax.in) break;                              // (block-uniform: an even w has a lone last chunk)
      const int j0 = c == 0 ? kRowF / 2 : 0;
      const int ncol = !active ? j0 : (c == ax.in ? kRowF / 2 : kRowF);
      const int xc = kRowF * c - kRowF / 2;              // x of the chunk's byte 0
      // classes in pairs for the packed fp32x2 pipe; an odd K is padded with a flat line far below every real logit.
      // dl: the slopes once more, dynamically indexed (local)
      constexpr int KP = (K + 1) / 2;
      float2 T2[KP], D2[KP];
      float dl[K];
      float M = 0.f;
      const float* sp = s_src + 2 * cc;
#pragma unroll
      for (int k = 0; k < 2 * KP; ++k) {
        float t = -1.0e30f, d = 0.f;
        if (k < K) {
          const float2 a = *reinterpret_cast<const float2*>(sp + k * (kRowSrc * 4) + ro0);
          const float2 bb = *reinterpret_cast<const float2*>(sp + k * (kRowSrc * 4) + ro1);
          const float2 tu = __ffma2_rn(hy0, a, __fmul2_rn(hy1, bb));     // (T, U): the vertical lerp at both columns
          t = tu.x;
          M = fmaxf(M, fmaxf(fabsf(tu.x), fabsf(tu.y)));
          d = __fsub_rn(tu.y, tu.x);
          dl[k] = d;
        }
        if (k & 1) { T2[k >> 1].y = t; D2[k >> 1].y = d; }
        else { T2[k >> 1].x = t; D2[k >> 1].x = d; }
      }
      const float gap = __fmaf_rn(M, 1.52587890625e-05f, kTieGap);       // 1e-5 + 2^-16 * M
      const bool sane = M < 1.0e29f;                     // (false for NaN / Inf / absurd logits: no fast path at all)
      uint32_t wd[kRowF / 4];
#pragma unroll
      for (int w = 0; w < kRowF / 4; ++w) wd[w] = 0u;
      uint32_t prev = 0u;
      int r = j0;
      bool last_unc = false;
      // warp-uniform trip count (every lane stays until the slowest row of the warp is done): the code behind the
      // loop then runs converged
      while (__any_sync(0xffffffffu, r < ncol)) {
        if (r >= ncol) continue;
        const float l = __fmul_rn((float)r + 0.5f, 1.f / kRowF);         // exact
        const float2 l2 = make_float2(l, l);
        // (the maxima and the candidate bits below are folded as TWO interleaved chains each: the kernel lives on
        //  latency — 5 warps per scheduler, issue slots half idle — and a 6-deep FMNMX3 / 12-deep SHF chain is most
        //  of a sweep step's critical path)
        float2 v2[KP];
        float m = -INFINITY, mb = -INFINITY;
#pragma unroll
        for (int j = 0; j < KP; ++j) {
          v2[j] = __ffma2_rn(l2, D2[j], T2[j]);
          if (j & 1) mb = fmaxf(mb, fmaxf(v2[j].x, v2[j].y));
          else m = fmaxf(m, fmaxf(v2[j].x, v2[j].y));
        }
        m = fmaxf(m, mb);
        const float thr = __fsub_rn(m, gap);
        const float2 thr2 = make_float2(thr, thr), neg1 = make_float2(-1.f, -1.f);
        // n_k = thr - v_k: negative exactly for the classes within the gap of the leader (sign bits funnel-shifted
        // together: bit 2*KP-1-k <=> class k), positive = the lead to lose
        float2 n2[KP];
        constexpr int KH = KP / 2;                       // pairs [0, KH) and [KH, KP) gather their bits side by side
        uint32_t cand = 0, candb = 0;
#pragma unroll
        for (int j = 0; j < KP; ++j) {
          n2[j] = __ffma2_rn(v2[j], neg1, thr2);
          if (j < KH) {
            cand = __funnelshift_l(__float_as_uint(n2[j].x), cand, 1);
            cand = __funnelshift_l(__float_as_uint(n2[j].y), cand, 1);
          } else {
            candb = __funnelshift_l(__float_as_uint(n2[j].x), candb, 1);
            candb = __funnelshift_l(__float_as_uint(n2[j].y), candb, 1);
          }
        }
        cand = (cand << (2 * (KP - KH))) | candb;
        const bool unc = !sane || cand == 0u || (cand & (cand - 1)) != 0;   // (cand == 0: NaN logits)
        uint32_t cls = 0u;
        int rend = r + 1;
        if (!unc) {
          const int a = __clz(cand) - (32 - 2 * KP);
          const float nDa = -dl[a];
          const float2 nDa2 = make_float2(nDa, nDa);
          float rmax = 0.f, rmaxb = 0.f;
#pragma unroll
          for (int j = 0; j < KP; ++j) {                 // class a itself: 0 * (1 / -gap) = -0, never the maximum
            const float2 e = __fadd2_rn(D2[j], nDa2);
            const float2 q = __fmul2_rn(e, make_float2(rcp_approx_row(n2[j].x), rcp_approx_row(n2[j].y)));
            if (j & 1) rmaxb = fmaxf(rmaxb, fmaxf(q.x, q.y));
            else rmax = fmaxf(rmax, fmaxf(q.x, q.y));
          }
          rmax = fmaxf(rmax, rmaxb);
          cls = (uint32_t)a;
          rend = ncol;
          if (rmax > 0.f) {
            const float step = __fmul_rn(rcp_approx_row(rmax), 0.999755859375f);   // (1 - 2^-12) / rmax
            const float hi = __fadd_rn(l, step);
            // the first column whose weight (j + 0.5) / 32 is >= hi (exact: x32 and -0.5 round nothing here)
            const int j = (int)fminf(ceilf(__fmaf_rn(hi, (float)kRowF, -0.5f)), 1.0e6f);
            rend = max(r + 1, min(j, ncol));
          }
        } else if (!enqueue(xc + r)) {
          ovf = true;
        }
        if (!(unc && last_unc)) {                        // a new run from byte r on (an uncertain stretch is class 0 for now)
          const uint32_t delta = (cls ^ prev) * 0x01010101u;
          prev = cls;
          const int r8 = 8 * r;
#pragma unroll
          for (int w = 0; w < kRowF / 4; ++w) wd[w] ^= ones_from_bit(max(r8 - 32 * w, 0)) & delta;
        }
        last_unc = unc;
        r = rend;
      }
      if (active) {
        uint4* o = reinterpret_cast<uint4*>(orow + xc);
        if (c != 0) o[0] = make_uint4(wd[0], wd[1], wd[2], wd[3]);
        if (c != ax.in) o[1] = make_uint4(wd[4], wd[5], wd[6], wd[7]);
        if (HIST) {
          const uint4* gp = reinterpret_cast<const uint4*>(grow + xc);
          if (c != 0) {
            const uint4 gw = __ldg(gp);
            h.word_trusted(wd[0], gw.x); h.word_trusted(wd[1], gw.y); h.word_trusted(wd[2], gw.z); h.word_trusted(wd[3], gw.w);
          }
          if (c != ax.in) {
            const uint4 gw = __ldg(gp + 1);
            h.word_trusted(wd[4], gw.x); h.word_trusted(wd[5], gw.y); h.word_trusted(wd[6], gw.z); h.word_trusted(wd[7], gw.w);
          }
        }
      }
    }

    // ---- cold epilogue: the warp's queued pixels, spread over its lanes, through the pinned softmax
    const bool warp_ovf = __any_sync(0xffffffffu, ovf);
    __syncwarp();                                        // queue and row stores were written by other lanes of this warp
    const int nq = warp_ovf ? 0 : min(s_qn[warp_in_block], kQueue);
    for (int e = (int)(threadIdx.x & 31); e < nq; e += 32) {
      const uint32_t ent = s_q[warp_in_block][e];
      const int yy = Y0 + (int)(ent >> 16), x = (int)(ent & 0xffffu);
      const TapH tx = tap(ax, x), tyy = tap(ay, yy);
      Vals<K> vals;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const float* r0 = lb + k * plane + tyy.i0 * ax.in;
        const float* r1 = lb + k * plane + tyy.i1 * ax.in;
        vals.v[k] = lerp2(tyy.l0, lerp2(tx.l0, __ldg(r0 + tx.i0), tx.l1, __ldg(r0 + tx.i1)), tyy.l1,
                          lerp2(tx.l0, __ldg(r1 + tx.i0), tx.l1, __ldg(r1 + tx.i1)));
      }
      const int cls = exact_from_values<K>(vals);
      const int64_t o = ((int64_t)b * ay.out + yy) * ax.out + x;
      mask[o] = (uint8_t)cls;
      if (HIST && cls != 0) h.move(gt[o], 0u, (uint32_t)cls);
    }
    if (warp_ovf && active) {
      // a queue overflowed (e.g. constant logits): every lane re-resolves its own 64 pixels with the exact rule
      const int xa = max(kRowF * 2 * p - kRowF / 2, 0), xb = min(kRowF * (2 * p + 2) - kRowF / 2, ax.out);
      for (int x = xa; x < xb; ++x) {
        const TapH tx = tap(ax, x);
        const uint32_t now = (uint32_t)exact_pixel(lb, K, plane, ax.in, ty.i0, ty.i1, ty.l0, ty.l1, tx.i0, tx.i1,
                                                   tx.l0, tx.l1);
        if (HIST) {
          const uint32_t was = orow[x];                  // (this thread's own store)
          if (now != was) h.move(grow[x], was, now);
        }
        orow[x] = (uint8_t)now;
      }
    }
  }
  if (HIST) {
    __syncthreads();
    h.flush(s_hist, C, status, true);
    if (PUSH) xchg_push_tail(C, px, s_step, gridDim.x);
  }
}

// Can the row form take this shape?  (x32 horizontal lift; a block's 256 rows within kRowSrc source rows; K <= 15:
// candidate bits; 16-byte stores)
bool lift_argmax_row_ok(int K, int h, int w, int H, int W, const void* mask) {
  return K >= 1 && K <= 15 && W == kRowF * w && W <= 65504 && (255ll * h + H - 1) / H + 3 <= kRowSrc &&
         (reinterpret_cast<uintptr_t>(mask) & 15u) == 0;
}

// gt == nullptr: mask only.  px != nullptr: histogram + peer push.
int launch_lift_argmax_row(const float* logits, uint8_t* mask, const uint8_t* gt, int64_t* C, int B, int K, int h,
                           int w, int H, int W, int* status, const XchgPush* px, cudaStream_t st) {
  AxisH ay{(float)h / (float)H, h, H}, ax{(float)w / (float)W, w, W};
  const int npairs = (w + 2) / 2;                        // chunks 0 .. w, two per thread
  const int nrb = (H + kRowThreads - 1) / kRowThreads;
  const int64_t items = (int64_t)npairs * nrb * B;
  if (items > 0x7fffffff) return LDIFF_EUNSUPPORTED;
  const int64_t cap = (int64_t)sm_count() * (2048 / kRowThreads);
  const int grid = (int)(items < cap ? items : cap);
  unsigned long long* Cu = reinterpret_cast<unsigned long long*>(C);
  switch (K) {
#define LR(KK) case KK:                                                                                             \
    if (px) lift_argmax_row_kernel<KK, true, true><<<grid, kRowThreads, 0, st>>>(logits, mask, gt, Cu, ay, ax, npairs, nrb, B, status, *px); \
    else if (gt) lift_argmax_row_kernel<KK, true, false><<<grid, kRowThreads, 0, st>>>(logits, mask, gt, Cu, ay, ax, npairs, nrb, B, status, XchgPush{}); \
    else lift_argmax_row_kernel<KK, false, false><<<grid, kRowThreads, 0, st>>>(logits, mask, nullptr, nullptr, ay, ax, npairs, nrb, B, status, XchgPush{}); \
    break;
    LR(1) LR(2) LR(3) LR(4) LR(5) LR(6) LR(7) LR(8) LR(9) LR(10) LR(11) LR(12) LR(13) LR(14) LR(15)
#undef LR
    default: return LDIFF_EUNSUPPORTED;
  }
  return check_launch();
}

}  // namespace ldiff
