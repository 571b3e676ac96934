// Shared pieces of the a-6 confusion histogram (sm_100a): the per-block shared-memory histogram that
// confusion.cu's stand-alone kernel and the fused producers (lut_paint_hist in confusion.cu,
// lift_argmax_hist in head.cu) all use, and the NVLink peer-window push that any of them can carry as its
// tail (multi-GPU: the histogram and its all-reduce are one kernel).
#pragma once
#include "common.cuh"

namespace ldiff {

// bytes of w that are >= k (k <= 128) get 0x80, others 0
__device__ __forceinline__ uint32_t bytes_ge(uint32_t w, uint32_t k) {
  return (((w & 0x7f7f7f7fu) + (0x80u - k) * 0x01010101u) | w) & 0x80808080u;
}

// ---- peer exchange window (multi-GPU): one per rank, mapped into every peer over NVLink -------------
// The only cross-rank step of the path is the SUM of the int64 matrices.  Instead of a separate
// collective, the LAST block of the kernel that finishes a matrix stores it straight into every
// peer's window (plain 8-byte stores over NVLink peer mappings); a one-block kernel on each rank then
// waits for the W rows and adds them.  Every 8-byte word carries half a counter and the step number
// (data and "it has landed" travel in one atomic store, as in NCCL's LL protocol), so the pusher
// needs no system fence and no separate flag: its tail is one dependent read of the matrix and a
// burst of fire-and-forget stores.  Rows are overwritten, never accumulated, so nothing is zeroed
// between steps; kXSlots ring slots keep a row alive until every rank has read it (ordering
// contract in ldiff.h).
constexpr int kXSlots = 4, kXMaxWorld = 16, kXMaxChan = 4, kXHeaderBytes = 256;
struct XchgHeader {
  unsigned long long step[kXMaxChan];                          // pushes completed by THIS rank, per channel
  unsigned long long reduced;                                  // reduces completed by THIS rank
  unsigned int ticket[kXMaxChan];                              // last-block election of the pushing grid
};
static_assert(sizeof(XchgHeader) <= kXHeaderBytes, "header does not fit");
struct XchgPush {                                              // by-value kernel argument; win == nullptr: off
  XchgHeader* win;
  int world, rank, channels, channel, n;
  unsigned long long peers[kXMaxWorld];                        // window base of every rank, as mapped here
};
// fills *px from an exchange handle (confusion.cu); LDIFF_EINVAL if the channel / matrix size do not fit
int xchg_push_args(void* handle, int channel, int n_i64, XchgPush* px);

// row of (slot, source rank, channel): 2*n words, word 2*bin = lo32 | tag<<32, word 2*bin+1 = hi32 | tag<<32
__device__ __forceinline__ unsigned long long* xchg_row(unsigned long long base, int slot, int src, int chan,
                                                        int world, int channels, int n) {
  return reinterpret_cast<unsigned long long*>(base + kXHeaderBytes) +
         ((int64_t)(slot * world + src) * channels + chan) * (2 * n);
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// (push) this launch's step number: only the previous launch's last block ever changes the counter, so a
// block reads it at its start, long before the tail needs it, instead of on the tail's dependent chain
__device__ __forceinline__ unsigned long long xchg_step_of_launch(const XchgPush& px) {
  return __ldcg(&px.win->step[px.channel]) + 1;
}

// Tail of a kernel whose blocks have just added their share into C with global atomics: the last block
// to arrive (ticket election behind a fence) owns the complete matrix and pushes it to every rank.
// Called by ALL threads of EVERY block (contains __syncthreads); nblocks = blocks of the whole grid.
__device__ __forceinline__ void xchg_push_tail(const unsigned long long* __restrict__ C, const XchgPush& px,
                                               unsigned long long step, unsigned int nblocks) {
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(&px.win->ticket[px.channel], 1u) == nblocks - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const unsigned long long tag = (step & 0xffffffffull) << 32;
  const int slot = (int)(step % kXSlots);
  for (int bin = threadIdx.x; bin < px.n; bin += blockDim.x) {
    const unsigned long long v = __ldcg(C + bin);
    const unsigned long long w0 = (v & 0xffffffffull) | tag, w1 = (v >> 32) | tag;
    for (int q = 0; q < px.world; ++q) {
      unsigned long long* row = xchg_row(px.peers[q], slot, px.rank, px.channel, px.world, px.channels, px.n);
      st_relaxed_sys(row + 2 * bin, w0);
      st_relaxed_sys(row + 2 * bin + 1, w1);
    }
  }
  if (threadIdx.x == 0) {
    px.win->step[px.channel] = step;
    px.win->ticket[px.channel] = 0;                    // re-armed for the next launch (stream-ordered)
  }
}

// ---- per-block histogram ---------------------------------------------------------------------------
// hist[bin][R] in shared memory, replicated per lane (bank == lane: a warp-wide update never bank-conflicts
// and never hits the same address twice); bin = gt * K + pred, gt clamped to K ("other" row).
// SMALLK (K <= 15): branch-free SWAR on 4 packed pixels, bins fit a byte.
template <int R, bool SMALLK>
struct BlockHist {
  uint32_t* my;            // this lane's replica column
  uint32_t K, k4, bad;

  __device__ __forceinline__ void init(uint32_t* hist, int K_) {     // caller: __syncthreads() afterwards
    const int nbins = (K_ + 1) * K_;
    for (int i = threadIdx.x; i < nbins * R; i += blockDim.x) hist[i] = 0;
    my = hist + (threadIdx.x & (R - 1));
    K = (uint32_t)K_;
    k4 = K * 0x01010101u;
    bad = 0;
  }
  __device__ __forceinline__ void pixel(uint32_t p, uint32_t g) {
    g = min(g, K);
    if (p >= K) { bad = 1; p = 0; }
    atomicAdd(my + (g * K + p) * R, 1u);
  }
  // move one pixel of ground truth g from predicted class p_old to p_new (both < K)
  __device__ __forceinline__ void move(uint32_t g, uint32_t p_old, uint32_t p_new) {
    g = min(g, K);
    atomicSub(my + (g * K + p_old) * R, 1u);
    atomicAdd(my + (g * K + p_new) * R, 1u);
  }
  __device__ __forceinline__ void word(uint32_t pw, uint32_t gw) {    // 4 packed pixels
    if (SMALLK) {
      const uint32_t gm = (bytes_ge(gw, K) >> 7) * 0xffu;             // 0xff where gt >= K
      const uint32_t gc = (gw & ~gm) | (k4 & gm);
      const uint32_t po = bytes_ge(pw, K);
      bad |= po;
      const uint32_t pc = pw & ~((po >> 7) * 0xffu);
      const uint32_t bins = gc * K + pc;                              // four byte-sized bin indices
      atomicAdd(my + (bins & 0xffu) * R, 1u);
      atomicAdd(my + ((bins >> 8) & 0xffu) * R, 1u);
      atomicAdd(my + ((bins >> 16) & 0xffu) * R, 1u);
      atomicAdd(my + (bins >> 24) * R, 1u);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) pixel((pw >> (8 * k)) & 0xffu, (gw >> (8 * k)) & 0xffu);
    }
  }
  // 4 packed pixels whose predictions are known to be < K (a fused producer that emits them itself): no range check
  __device__ __forceinline__ void word_trusted(uint32_t pw, uint32_t gw) {
    static_assert(SMALLK, "byte-sized bins only");
    const uint32_t gm = (bytes_ge(gw, K) >> 7) * 0xffu;               // 0xff where gt >= K
    const uint32_t bins = ((gw & ~gm) | (k4 & gm)) * K + pw;
    atomicAdd(my + (bins & 0xffu) * R, 1u);
    atomicAdd(my + ((bins >> 8) & 0xffu) * R, 1u);
    atomicAdd(my + ((bins >> 16) & 0xffu) * R, 1u);
    atomicAdd(my + (bins >> 24) * R, 1u);
  }
  // after a __syncthreads(): fold the replicas, one 64-bit global atomic per non-empty bin.  wrap32: the
  // block used move(), so a replica may hold a count "below zero" that another replica compensates; the
  // fold is then taken modulo 2^32 (a block never owns 2^32 pixels of one bin in that mode)
  __device__ __forceinline__ void flush(const uint32_t* hist, unsigned long long* __restrict__ C,
                                        int* __restrict__ status, bool wrap32 = false) {
    if (bad) atomicOr(status, LDIFF_STATUS_PRED_RANGE);
    const int nbins = (int)((K + 1) * K);
    for (int bin = threadIdx.x; bin < nbins; bin += blockDim.x) {
      unsigned long long s = 0;
#pragma unroll 8
      for (int r = 0; r < R; ++r) s += hist[bin * R + ((r + threadIdx.x) & (R - 1))];
      if (wrap32) s &= 0xffffffffull;
      if (s) atomicAdd(C + bin, s);
    }
  }
};

}  // namespace ldiff
