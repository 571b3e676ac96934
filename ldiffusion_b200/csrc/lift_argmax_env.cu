// a-5 lift + argmax, envelope form (COLUMN form: a thread sweeps output columns; the x32 lifts of the path run the
// row form, lift_argmax_row.cu, and come here for other factors), with the a-6 confusion histogram optionally fused in (sm_100a).
#include "head_common.cuh"
#include "hist.cuh"

namespace ldiff {

// ----------------------------------------------------------------------------
// Envelope form: per output column the K lifted logits of a band are K LINES in the
// vertical weight l — v_k(l) = T_k + l * (U_k - T_k) — so the argmax over the band's rows is the upper
// envelope of K lines: a handful of intervals (2.4 on i.i.d. logits, 1 on smooth maps), not K values
// per pixel.  A thread sweeps its column once: at the current row it evaluates the K lines (as the
// kernel above does for EVERY row), and if one class leads all others by more than the near-tie gap it
// solves for how far down the band that lead is guaranteed to last:
//     lead over class k at l' : (m - v_k) - (l' - l) * (D_k - D_a)  >  gap        (linear in l')
//     =>  l' - l  <  1 / max_k [ (D_k - D_a) / ((m - gap) - v_k) ]                 (one MUFU.RCP per class)
// writes that run as ONE segment (start row, class) and jumps to the first row behind it.  Rows where
// the lead is within the gap become single-row "uncertain" segments and are queued for the pinned softmax
// exactly as before.  The segments of a column (<= 7, packed in two 64-bit byte vectors) are then replayed
// by a row loop that costs 4 instructions per pixel (ISETP, IADD, 2 PRMT) + the packed store: ~69 -> ~20
// instructions per pixel on i.i.d. logits.
//
// Why a run is safe (M = the column's largest |T_k|, |U_k|; w_k(l) the exact line through the fp32 T_k, U_k):
//  * the reference's v_ref = fma(fl(1-l), T, fl(l*U)) is within 2.5 * 2^-24 * M of w_k(l);
//  * the swept values v' = fma(l, fl(U-T), T) are within 4 * 2^-24 * M of w_k(l), the slope differences
//    fl(D_k - D_a) within 8 * 2^-24 * M, so a lead extrapolated over l' - l <= 1 is off by < 2^-19 * M;
//  * the approximate reciprocals / products put a relative 2^-21 on l' - l, i.e. < 2^-19 * M on the lead,
//    and the run is shortened by another 2^-12 relative for good measure;
//  so with gap = 1e-5 + 2^-17 * M every pixel of a run has a reference top-2 gap above 1e-5 and the same
//  winner, and argmax(softmax) == argmax there (oracle/head.py).  Everything else goes to the pinned softmax,
//  which recomputes the reference's own roundings from the logits: the fast path only decides WHO is safe.
//
// HIST: the a-6 confusion histogram of the mask against gt is accumulated in the same pass (the packed
// class word of the row loop meets the gt word in registers; uncertain pixels are entered as class 0 and
// moved to their resolved class by the epilogue), PUSH: with the NVLink peer push as the kernel's tail.
constexpr int kEnvSegs = 7;        // segments per column and band (byte 7 of the packed vectors stays 0)
constexpr int kEnvSrcW = 40;       // source columns a block stages in shared memory
constexpr int kEnvRows = 127;
constexpr int kEnvThreads = 256;   // threads per block.  (128 measured 22.8 / 24.2 us on i.i.d. / smooth logits against
                                   // 21.6 / 17.2 us with 256: the per-block prologue is paid twice as often)      // rows per band (start rows are bytes; 0xff = "no further segment")

// byte `sel & 7` of {hi, lo} in byte 0 (selector nibbles 1..3 = 7: the always-zero top byte of hi) — the raw PRMT;
// __byte_perm would mask the selector first
__device__ __forceinline__ uint32_t pick_byte(uint32_t lo, uint32_t hi, uint32_t sel) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(lo), "r"(hi), "r"(sel));
  return r;
}

__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));   // one MUFU (no denormal bracket)
  return r;
}

template <int K, int COLS, bool HIST, bool PUSH>
__global__ void __launch_bounds__(kEnvThreads, 1024 / kEnvThreads)   // 64 registers
lift_argmax_env_kernel(const float* __restrict__ logits, uint8_t* __restrict__ mask, const uint8_t* __restrict__ gt,
                       unsigned long long* __restrict__ C, AxisH ay, AxisH ax, int nxb, int B,
                       int* __restrict__ status, XchgPush px) {
  __shared__ float2 s_l[kBand];
  __shared__ float s_src[2 * K * kEnvSrcW];
  __shared__ uint32_t s_q[kEnvThreads / 32][kQueue];
  __shared__ int s_qn[kEnvThreads / 32];
  __shared__ int s_y[2];
  __shared__ uint32_t s_hist[HIST ? (K + 1) * K * 32 : 1];
  // the band's ground-truth tile, fetched with cp.async at the start of the item and read by the row loop long
  // after it has landed (a global load per row put ~7 us of exposed latency into the histogram form)
  constexpr int kGtRows = (HIST && K <= 12) ? 32 : 1;    // (the 1.5x taller first band reads gt from global)
  constexpr int kRowBytes = kEnvThreads * COLS;
  __shared__ __align__(16) uint8_t s_gt[kGtRows * kRowBytes];
  __shared__ unsigned long long s_step;
  BlockHist<32, true> h;
  if (HIST) h.init(s_hist, K);
  if (PUSH && threadIdx.x == 0) s_step = xchg_step_of_launch(px);
  const int warp_in_block = threadIdx.x >> 5;
  const int plane = ay.in * ax.in;
  const float inv_scale = 1.f / ay.scale;
  auto first_row = [&](int i) {
    if (i <= 0) return 0;
    int y = (int)ceilf(((float)i + 0.5f) / ay.scale - 0.5f);
    y = max(0, min(y, ay.out));
    while (y > 0 && tap(ay, y - 1).i0 >= i) --y;
    while (y < ay.out && tap(ay, y).i0 < i) ++y;
    return y;
  };
  const int items = nxb * ay.in * B;
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int bx = item % nxb, rest = item / nxb;
    const int iy = rest % ay.in, b = rest / ay.in;
    __syncthreads();                                     // previous item's epilogue has read s_l / s_src / s_q / s_y
    if (threadIdx.x < 2)                                 // the band's first and one-past-last output row
      s_y[threadIdx.x] = (threadIdx.x == 1 && iy + 1 >= ay.in) ? ay.out : first_row(iy + (int)threadIdx.x);
    if (threadIdx.x < kEnvThreads / 32) s_qn[threadIdx.x] = 0;
    __syncthreads();
    const int Y0 = s_y[0], Y1 = s_y[1];
    if (Y0 >= Y1) continue;                              // (block-uniform)
    for (int i = threadIdx.x; i < Y1 - Y0; i += kEnvThreads) {
      const TapH t = tap(ay, Y0 + i);
      s_l[i] = make_float2(t.l0, t.l1);
    }
    bool gt_staged = false;
    if (HIST && kGtRows > 1) {
      const uint8_t* g0 = gt + ((int64_t)b * ay.out + Y0) * ax.out + bx * kRowBytes;
      gt_staged = (Y1 - Y0) <= kGtRows && (kRowBytes % 16) == 0 && (ax.out % 16) == 0 &&
                  (reinterpret_cast<uintptr_t>(gt) & 15u) == 0 && (bx + 1) * kRowBytes <= ax.out;
      if (gt_staged) {
        constexpr int kChunks = kRowBytes / 16 > 0 ? kRowBytes / 16 : 1;
        for (int i = threadIdx.x; i < (Y1 - Y0) * kChunks; i += kEnvThreads) {
          const int r = i / kChunks, ch = i - r * kChunks;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(
                           s_gt + r * kRowBytes + ch * 16)), "l"(g0 + (int64_t)r * ax.out + ch * 16) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
    }
    const int i0 = min(iy, ay.in - 1), i1 = min(iy + 1, ay.in - 1);   // the band's source row pair
    const float* lb = logits + (int64_t)b * K * plane;
    // stage the two source rows of every class for this block's columns: warp w takes (row, class) pairs
    // w, w + 8, ..., lane = source column (no index divisions)
    const int xfirst = bx * kEnvThreads * COLS, xlast = min(xfirst + kEnvThreads * COLS, ax.out) - 1;
    const int c_lo = tap(ax, xfirst).i0, c_hi = tap(ax, xlast).i1;
    const int ncol = c_hi - c_lo + 1;
    const bool staged = ncol <= kEnvSrcW;
    if (staged) {
      for (int rk = warp_in_block; rk < 2 * K; rk += kEnvThreads / 32) {
        const int rs = rk >= K ? 1 : 0, k = rk - rs * K;
        const float* row = lb + k * plane + (rs ? i1 : i0) * ax.in + c_lo;
        for (int j = (int)(threadIdx.x & 31); j < ncol; j += 32) s_src[rk * ncol + j] = __ldg(row + j);
      }
    }
    const int kstride = staged ? ncol : plane;
    const int rowoff = staged ? K * ncol : (i1 - i0) * ax.in;
    const int base0 = staged ? -c_lo : i0 * ax.in;
    __syncthreads();

    const int x0 = xfirst + (int)threadIdx.x * COLS;
    const bool active = x0 < ax.out;                     // (W % COLS == 0: a thread's columns are all in or all out)
    const int nrow = active ? Y1 - Y0 : 0;
    bool ovf = false;
    auto enqueue = [&](int r, int c) -> bool {           // false: the warp's queue is full
      const int slot = atomicAdd(&s_qn[warp_in_block], 1);
      if (slot >= kQueue) return false;
      s_q[warp_in_block][slot] = ((uint32_t)r << 16) | (uint32_t)(threadIdx.x * COLS + c);
      return true;
    };

    uint32_t slo[COLS], shi[COLS], clo[COLS], chi[COLS];
#pragma unroll
    for (int c = 0; c < COLS; ++c) {
      slo[c] = 0xffffffffu; shi[c] = 0x00ffffffu; clo[c] = 0u; chi[c] = 0u;
      {                                                  // (every lane, active or not: the sweep loop votes warp-wide)
        const TapH tx = tap(ax, min(x0 + c, ax.out - 1));
        // classes in pairs for the packed fp32x2 pipe (FFMA2 / FADD2 / FMUL2: IEEE per lane); an odd K is padded
        // with a flat line far below every real logit.  dl: the slopes once more, dynamically indexed (local)
        constexpr int KP = (K + 1) / 2;
        float2 T2[KP], D2[KP];
        float dl[K];
        float M = 0.f;
        auto column = [&](auto src_at) {                 // horizontal lift of the two source rows, once per band
#pragma unroll
          for (int k = 0; k < 2 * KP; ++k) {
            float t = -1.0e30f, d = 0.f;
            if (k < K) {
              const int o0 = base0 + k * kstride + tx.i0, o1 = base0 + k * kstride + tx.i1;
              t = lerp2(tx.l0, src_at(o0), tx.l1, src_at(o1));
              const float u = lerp2(tx.l0, src_at(o0 + rowoff), tx.l1, src_at(o1 + rowoff));
              M = fmaxf(M, fmaxf(fabsf(t), fabsf(u)));
              d = __fsub_rn(u, t);
              dl[k] = d;
            }
            if (k & 1) { T2[k >> 1].y = t; D2[k >> 1].y = d; }
            else { T2[k >> 1].x = t; D2[k >> 1].x = d; }
          }
        };
        if (staged) column([&](int i) { return s_src[i]; });            // (block-uniform)
        else column([&](int i) { return __ldg(lb + i); });
        const float gap = __fmaf_rn(M, 7.62939453125e-06f, kTieGap);        // 1e-5 + 2^-17 * M
        const bool sane = M < 1.0e29f;                   // (false for NaN / Inf / absurd logits: no fast path at all)
        int nseg = 0, r = 0;
        bool last_unc = false;
        // warp-uniform trip count (every lane stays until the slowest column of the warp is done): the code behind
        // the loop then runs converged — a per-lane `while (r < nrow)` left the warp split for the rest of the kernel
        while (__any_sync(0xffffffffu, r < nrow)) {
          if (r >= nrow) continue;
          const float l = s_l[r].y;
          const float2 l2 = make_float2(l, l);
          float2 v2[KP];
          float m = -INFINITY;
#pragma unroll
          for (int j = 0; j < KP; ++j) {
            v2[j] = __ffma2_rn(l2, D2[j], T2[j]);
            m = fmaxf(m, fmaxf(v2[j].x, v2[j].y));
          }
          const float thr = __fsub_rn(m, gap);
          const float2 thr2 = make_float2(thr, thr), neg1 = make_float2(-1.f, -1.f);
          // n_k = thr - v_k: negative exactly for the classes within the gap of the leader (its sign bits, funnel-
          // shifted together, are the candidate mask: bit 2*KP-1-k <=> class k), positive = the lead to lose
          float2 n2[KP];
          uint32_t cand = 0;
#pragma unroll
          for (int j = 0; j < KP; ++j) {
            n2[j] = __ffma2_rn(v2[j], neg1, thr2);
            cand = __funnelshift_l(__float_as_uint(n2[j].x), cand, 1);
            cand = __funnelshift_l(__float_as_uint(n2[j].y), cand, 1);
          }
          cand &= (1u << (2 * KP)) - 1u;
          bool unc = !sane || cand == 0u || (cand & (cand - 1)) != 0;   // (cand == 0: NaN logits)
          int cls = 0, rend = r + 1;
          if (!unc) {
            const int a = __clz(cand) - (32 - 2 * KP);
            const float nDa = -dl[a];
            const float2 nDa2 = make_float2(nDa, nDa);
            float rmax = 0.f;
#pragma unroll
            for (int j = 0; j < KP; ++j) {               // class a itself: 0 * (1 / -gap) = -0, never the maximum
              const float2 e = __fadd2_rn(D2[j], nDa2);
              const float2 q = __fmul2_rn(e, make_float2(rcp_approx(n2[j].x), rcp_approx(n2[j].y)));
              rmax = fmaxf(rmax, fmaxf(q.x, q.y));
            }
            cls = a;
            rend = nrow;
            if (rmax > 0.f) {
              const float step = __fmul_rn(rcp_approx(rmax), 0.999755859375f);   // (1 - 2^-12) / rmax
              const float hi = __fadd_rn(l, step);
              int j = r + (int)fminf(ceilf(__fmul_rn(step, inv_scale)), 1.0e6f);
              j = max(r + 1, min(j, nrow));
              while (j > r + 1 && s_l[j - 1].y >= hi) --j;     // rows r+1 .. j-1 all have l < hi (l is monotone)
              while (j < nrow && s_l[j].y < hi) ++j;
              rend = j;
            }
          }
          const bool newseg = !(unc && last_unc);
          if (newseg && nseg == kEnvSegs - 1 && (unc || rend < nrow)) {
            unc = true; cls = 0; rend = nrow;            // last slot: the rest of the column goes to the exact resolver
          }
          if (unc)
            for (int rr = r; rr < rend; ++rr)
              if (!enqueue(rr, c)) ovf = true;
          if (newseg) {
            if (nseg > 0) {                              // start row of segment nseg -> byte nseg-1 of the start vector
              const int sh = 8 * ((nseg - 1) & 3);
              const uint32_t flip = (0xffu ^ (uint32_t)r) << sh;
              if (nseg - 1 < 4) slo[c] ^= flip; else shi[c] ^= flip;
            }
            const uint32_t cw = (uint32_t)cls << (8 * (nseg & 3));
            if (nseg < 4) clo[c] |= cw; else chi[c] |= cw;
            ++nseg;
          }
          last_unc = unc;
          r = rend;
        }
      }
    }

    // ---- row loop: replay the segments; 4 instructions per pixel + the packed store (+ histogram)
    if (HIST && kGtRows > 1) {                           // (block-uniform)
      asm volatile("cp.async.wait_all;" ::: "memory");
      __syncthreads();
    }
    {
      uint32_t sel[COLS], nxt[COLS], cur[COLS];
#pragma unroll
      for (int c = 0; c < COLS; ++c) {
        sel[c] = 0x7770u;
        nxt[c] = pick_byte(slo[c], shi[c], sel[c]);
        cur[c] = pick_byte(clo[c], chi[c], sel[c]);
      }
      uint8_t* op = mask + ((int64_t)b * ay.out + Y0) * ax.out + x0;
      const uint8_t* gp = HIST ? (gt_staged ? s_gt + threadIdx.x * COLS : gt + ((int64_t)b * ay.out + Y0) * ax.out + x0)
                               : nullptr;
      const int gstep = gt_staged ? kRowBytes : ax.out;
      // 4 / COLS rows per iteration: their packed classes and gt bytes form one 4-pixel histogram word
      constexpr int RPW = 4 / COLS;
      int r = 0;
      for (; r + RPW <= nrow; r += RPW) {
        uint32_t pk[RPW], gk[RPW];
#pragma unroll
        for (int q = 0; q < RPW; ++q) {
          uint32_t packed = 0;
#pragma unroll
          for (int c = 0; c < COLS; ++c) {
            if ((uint32_t)(r + q) >= nxt[c]) ++sel[c];
            nxt[c] = pick_byte(slo[c], shi[c], sel[c]);
            cur[c] = pick_byte(clo[c], chi[c], sel[c]);
            packed |= cur[c] << (8 * c);
          }
          pk[q] = packed;
          gk[q] = 0;
          if (COLS == 1) { op[0] = (uint8_t)packed; if (HIST) gk[q] = gp[0]; }
          else if (COLS == 2) { *reinterpret_cast<uint16_t*>(op) = (uint16_t)packed; if (HIST) gk[q] = *reinterpret_cast<const uint16_t*>(gp); }
          else { *reinterpret_cast<uint32_t*>(op) = packed; if (HIST) gk[q] = *reinterpret_cast<const uint32_t*>(gp); }
          op += ax.out;
          if (HIST) gp += gstep;
        }
        if (HIST) {
          if (COLS == 4) h.word_trusted(pk[0], gk[0]);
          else if (COLS == 2) h.word_trusted(__byte_perm(pk[0], pk[RPW - 1], 0x5410), __byte_perm(gk[0], gk[RPW - 1], 0x5410));
          else h.word_trusted(__byte_perm(__byte_perm(pk[0], pk[1 % RPW], 0x0040), __byte_perm(pk[2 % RPW], pk[3 % RPW], 0x0040), 0x5410),
                              __byte_perm(__byte_perm(gk[0], gk[1 % RPW], 0x0040), __byte_perm(gk[2 % RPW], gk[3 % RPW], 0x0040), 0x5410));
        }
      }
      for (; r < nrow; ++r, op += ax.out) {              // the band's last (nrow mod RPW) rows
        uint32_t packed = 0;
#pragma unroll
        for (int c = 0; c < COLS; ++c) {
          if ((uint32_t)r >= nxt[c]) ++sel[c];
          nxt[c] = pick_byte(slo[c], shi[c], sel[c]);
          cur[c] = pick_byte(clo[c], chi[c], sel[c]);
          packed |= cur[c] << (8 * c);
        }
#pragma unroll
        for (int c = 0; c < COLS; ++c) {
          op[c] = (uint8_t)(packed >> (8 * c));
          if (HIST) h.pixel((packed >> (8 * c)) & 0xffu, gp[c]);
        }
        if (HIST) gp += gstep;
      }
    }

    // ---- cold epilogue: the warp's queued pixels, spread over its lanes, through the pinned softmax
    const bool warp_ovf = __any_sync(0xffffffffu, ovf);
    __syncwarp();                                        // the queue was written by other lanes of this warp
    const int nq = warp_ovf ? 0 : min(s_qn[warp_in_block], kQueue);
    const int xblock = xfirst;
    for (int e = (int)(threadIdx.x & 31); e < nq; e += 32) {
      const uint32_t ent = s_q[warp_in_block][e];
      const int r = (int)(ent >> 16), x = xblock + (int)(ent & 0xffffu);
      const TapH t = tap(ax, x);
      const float l0 = s_l[r].x, l1 = s_l[r].y;
      Vals<K> vals;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const float* r0 = lb + k * plane + i0 * ax.in;
        const float* r1 = lb + k * plane + i1 * ax.in;
        vals.v[k] = lerp2(l0, lerp2(t.l0, __ldg(r0 + t.i0), t.l1, __ldg(r0 + t.i1)), l1,
                          lerp2(t.l0, __ldg(r1 + t.i0), t.l1, __ldg(r1 + t.i1)));
      }
      const int cls = exact_from_values<K>(vals);
      const int64_t o = ((int64_t)b * ay.out + Y0 + r) * ax.out + x;
      mask[o] = (uint8_t)cls;
      if (HIST && cls != 0) h.move(gt[o], 0u, (uint32_t)cls);
    }
    if (warp_ovf) {
      // a queue overflowed (e.g. constant logits): every lane re-resolves its whole column(s) with the exact rule
      // and moves the pixels whose class differs from what the row loop entered
      for (int c = 0; c < COLS; ++c) {
        const TapH tx = tap(ax, min(x0 + c, ax.out - 1));
        uint32_t sel = 0x7770u;
        for (int rr = 0; rr < nrow; ++rr) {
          if ((uint32_t)rr >= __byte_perm(slo[c], shi[c], sel)) ++sel;
          const uint32_t was = __byte_perm(clo[c], chi[c], sel);
          const uint32_t now = (uint32_t)exact_pixel(lb, K, plane, ax.in, i0, i1, s_l[rr].x, s_l[rr].y, tx.i0, tx.i1,
                                                     tx.l0, tx.l1);
          if (now != was) {
            const int64_t o = ((int64_t)b * ay.out + Y0 + rr) * ax.out + x0 + c;
            mask[o] = (uint8_t)now;
            if (HIST) h.move(gt[o], was, now);
          }
        }
      }
    }
  }
  if (HIST) {
    __syncthreads();
    h.flush(s_hist, C, status, true);
    if (PUSH) xchg_push_tail(C, px, s_step, gridDim.x);
  }
}


template <int K, int COLS>
static void launch_env_k(const float* logits, uint8_t* mask, const uint8_t* gt, unsigned long long* C, AxisH ay,
                         AxisH ax, int nxb, int B, int grid, int* status, const XchgPush* px, cudaStream_t st) {
  if (px)
    lift_argmax_env_kernel<K, COLS, true, true><<<grid, kEnvThreads, 0, st>>>(logits, mask, gt, C, ay, ax, nxb, B, status, *px);
  else if (gt)
    lift_argmax_env_kernel<K, COLS, true, false><<<grid, kEnvThreads, 0, st>>>(logits, mask, gt, C, ay, ax, nxb, B, status,
                                                                       XchgPush{});
  else
    lift_argmax_env_kernel<K, COLS, false, false><<<grid, kEnvThreads, 0, st>>>(logits, mask, nullptr, nullptr, ay, ax, nxb, B,
                                                                        status, XchgPush{});
}

// Can the envelope kernel take this shape?  (K <= 15: candidate bits / byte-sized bins; bands of <= 127 rows)
bool lift_argmax_env_ok(int K, int h, int H) {
  return K >= 1 && K <= 15 && H >= h && max_band_rows(h, H) <= kEnvRows && h <= 65535;
}

// gt == nullptr: mask only.  px != nullptr: histogram + peer push.
int launch_lift_argmax_env(const float* logits, uint8_t* mask, const uint8_t* gt, int64_t* C, int B, int K, int h,
                           int w, int H, int W, int* status, const XchgPush* px, cudaStream_t st) {
  AxisH ay{(float)h / (float)H, h, H}, ax{(float)w / (float)W, w, W};
  // two columns per thread (16-bit stores) when the rows allow it.  (Four columns per thread measured slower:
  // 30.3 vs 24.8 us at the bench shape — half the warps, and the kernel lives on latency hiding.)
  const bool two = (W % 2) == 0 && ((reinterpret_cast<uintptr_t>(mask) | reinterpret_cast<uintptr_t>(gt)) & 1u) == 0;
  const int cols = two ? 2 : 1;
  const int nxb = (W / cols + kEnvThreads - 1) / kEnvThreads;
  const int64_t items = (int64_t)nxb * h * B;
  if (items > 0x7fffffff) return LDIFF_EUNSUPPORTED;
  const int64_t cap = (int64_t)sm_count() * (2048 / kEnvThreads);   // whole bands per block; a block walks several on big batches
  const int grid = (int)(items < cap ? items : cap);
  unsigned long long* Cu = reinterpret_cast<unsigned long long*>(C);
  switch (K) {
#define LE(KK) case KK:                                                                                      \
    if (two) launch_env_k<KK, 2>(logits, mask, gt, Cu, ay, ax, nxb, B, grid, status, px, st);                 \
    else launch_env_k<KK, 1>(logits, mask, gt, Cu, ay, ax, nxb, B, grid, status, px, st);                     \
    break;
    LE(1) LE(2) LE(3) LE(4) LE(5) LE(6) LE(7) LE(8) LE(9) LE(10) LE(11) LE(12) LE(13) LE(14) LE(15)
#undef LE
    default: return LDIFF_EUNSUPPORTED;
  }
  return check_launch();
}

}  // namespace ldiff
