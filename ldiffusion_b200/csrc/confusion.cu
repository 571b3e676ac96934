// a-6 confusion matrix C[(K+1),K] from uint8 prediction / ground-truth maps (sm_100a).
//
// One pass, 2 bytes per pixel.  The budget at the HBM roofline is ~11 issued
// instructions per pixel, so the per-pixel work is SWAR on the loaded words:
//  * 16 pixels per thread per iteration (two 128-bit streaming loads);
//  * per 32-bit word (4 pixels): branch-free bytewise "gt >= K -> K" clamp and
//    "pred >= K" check, then ONE multiply-add gives the four bin indices
//    gt*K + pred packed in bytes (bins fit a byte for K <= 15);
//  * one fire-and-forget shared-memory atomic per pixel into a histogram
//    replicated per lane (hist[bin][lane]: bank == lane, so a warp-wide update
//    never bank-conflicts and never hits the same address twice);
//  * at the end the block folds the 32 replicas and issues one 64-bit global
//    atomic per non-empty bin into the caller's int64 matrix (which may be the
//    buffer handed to ncclAllReduce).
// Larger K (16..128) takes the same kernel with per-pixel integer bin math.
#include <cstdlib>
#include <cstring>

#include "hist.cuh"

namespace ldiff {

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

template <int R, bool SMALLK, bool HAS_LUT, bool PUSH>
__device__ __forceinline__ void
confusion_hist_body(const uint8_t* __restrict__ pred, const uint8_t* __restrict__ gt,
                    const uint8_t* __restrict__ gt_lut, unsigned long long* __restrict__ C,
                    int64_t n, int K, int* __restrict__ status, XchgPush px) {
  extern __shared__ uint32_t hist[];                 // [nbins][R]
  __shared__ uint8_t lut[256];
  const int nbins = (K + 1) * K;
  BlockHist<R, SMALLK> h;
  h.init(hist, K);
  if (HAS_LUT)
    for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = gt_lut[i];
  __shared__ unsigned long long s_step;
  if (PUSH && threadIdx.x == 0) s_step = xchg_step_of_launch(px);
  __syncthreads();

  // grid.y = image (batched form: one matrix per image)
  pred += (int64_t)blockIdx.y * n;
  gt += (int64_t)blockIdx.y * n;
  C += (int64_t)blockIdx.y * nbins;

  auto pixel = [&](uint32_t p, uint32_t g) { h.pixel(p, HAS_LUT ? (uint32_t)lut[g] : g); };
  auto word = [&](uint32_t pw, uint32_t gw) {
    if (HAS_LUT)
      gw = (uint32_t)lut[gw & 0xff] | ((uint32_t)lut[(gw >> 8) & 0xff] << 8) |
           ((uint32_t)lut[(gw >> 16) & 0xff] << 16) | ((uint32_t)lut[gw >> 24] << 24);
    h.word(pw, gw);
  };

  // [0,head) scalar until both streams are 16-byte aligned (all of it if they are
  // aligned differently), then 16-pixel vectors, then the scalar tail
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t mp = (uint32_t)(reinterpret_cast<uintptr_t>(pred) & 15u);
  const uint32_t mg = (uint32_t)(reinterpret_cast<uintptr_t>(gt) & 15u);
  int64_t head = (mp == mg) ? (int64_t)((16u - mp) & 15u) : n;
  if (head > n) head = n;
  const int64_t nvec = (n - head) >> 4;
  for (int64_t i = gtid; i < head; i += stride) pixel(pred[i], gt[i]);
  const uint4* pv = reinterpret_cast<const uint4*>(pred + head);
  const uint4* gv = reinterpret_cast<const uint4*>(gt + head);
  for (int64_t v = gtid; v < nvec; v += stride) {
    const uint4 a = __ldcs(pv + v);
    const uint4 b = __ldcs(gv + v);
    word(a.x, b.x); word(a.y, b.y); word(a.z, b.z); word(a.w, b.w);
  }
  for (int64_t i = head + (nvec << 4) + gtid; i < n; i += stride) pixel(pred[i], gt[i]);
  __syncthreads();
  h.flush(hist, C, status);
  // ---- fused exchange: the last block to finish owns the complete matrix and pushes it to every rank
  if (PUSH) xchg_push_tail(C, px, s_step, gridDim.x * gridDim.y);
}

template <int R, bool SMALLK, bool HAS_LUT>
__global__ void __launch_bounds__(512)
confusion_hist_kernel(const uint8_t* __restrict__ pred, const uint8_t* __restrict__ gt,
                      const uint8_t* __restrict__ gt_lut, unsigned long long* __restrict__ C,
                      int64_t n, int K, int* __restrict__ status) {
  confusion_hist_body<R, SMALLK, HAS_LUT, false>(pred, gt, gt_lut, C, n, K, status, XchgPush{});
}

// same histogram with the fused push (its own entry so the plain kernel's register allocation
// stays what it was: the longer tail otherwise makes ptxas spill in the hot loop)
template <int R, bool SMALLK, bool HAS_LUT>
__global__ void __launch_bounds__(512, 2)
confusion_hist_push_kernel(const uint8_t* __restrict__ pred, const uint8_t* __restrict__ gt,
                           const uint8_t* __restrict__ gt_lut, unsigned long long* __restrict__ C,
                           int64_t n, int K, int* __restrict__ status, XchgPush px) {
  confusion_hist_body<R, SMALLK, HAS_LUT, true>(pred, gt, gt_lut, C, n, K, status, px);
}

// ----------------------------------------------------------------------------
// a-5 (cell form) + a-6 in one pass: mask[b,p] = lut[b][inst[b,p]] (conductor.py:224-231 + segmentor.py:536)
// AND C[gt[b,p]][mask[b,p]] += 1 while the mask byte is still in a register — the stand-alone histogram
// would read the 1 B/pixel mask back and needs its own launch.  6 B/pixel: 4 (instance id) + 1 (gt) in,
// 1 (mask) out.  K <= 15 (byte-sized bins); one matrix for the whole batch (grid.y = image picks the LUT).
// (256-thread blocks of <= 48 registers: 12 K registers and 17 KB of shared memory per block, so that blocks slot in
// next to the resident decode-tail CTAs of the concurrent pass instead of waiting for half an SM's register file)
template <bool PUSH, typename IdT>
__global__ void __launch_bounds__(256, 5)
lut_paint_hist_kernel(const IdT* __restrict__ inst, const uint8_t* __restrict__ lut, uint8_t* __restrict__ mask,
                      const uint8_t* __restrict__ gt, unsigned long long* __restrict__ C, int64_t n, int lut_size,
                      int64_t lut_stride, int K, int* __restrict__ status, XchgPush px) {
  extern __shared__ uint32_t hist[];                 // [nbins][32]
  BlockHist<32, true> h;
  h.init(hist, K);
  __shared__ unsigned long long s_step;
  if (PUSH && threadIdx.x == 0) s_step = xchg_step_of_launch(px);
  __syncthreads();
  const int b = blockIdx.y;
  // (global LUT through L1: beats a shared-memory copy at 800 entries — byte-wide 10.5 vs 9.4 us in round 1, and a
  //  nibble-packed shared copy, 100 words, measured 16.0 vs 13.1 us for this kernel on random ids: bank conflicts)
  // (ptxas re-derives lut + b * lut_stride in front of every byte lookup, 5 instructions; pinning the base in a
  //  register pair halves the kernel's address arithmetic — and measured +3..6 us per PASS in a same-box A/B
  //  (tools/ab_pass.sh): the kernel then issues its scattered loads in denser bursts next to the decode tails)
  const uint8_t* l = lut + b * lut_stride;
  const IdT* in = inst + b * n;
  const uint8_t* g = gt + b * n;
  uint8_t* out = mask + b * n;
  int badid = 0;
  auto look4 = [&](int4 q) -> uint32_t {
    const uint32_t a = (uint32_t)q.x, bb = (uint32_t)q.y, c = (uint32_t)q.z, d = (uint32_t)q.w;
    if (max(max(a, bb), max(c, d)) < (uint32_t)lut_size)          // common case: all four in range
      return (uint32_t)l[a] | ((uint32_t)l[bb] << 8) | ((uint32_t)l[c] << 16) | ((uint32_t)l[d] << 24);
    badid = 1;
    uint32_t w = 0;
    if (a < (uint32_t)lut_size) w |= l[a];
    if (bb < (uint32_t)lut_size) w |= (uint32_t)l[bb] << 8;
    if (c < (uint32_t)lut_size) w |= (uint32_t)l[c] << 16;
    if (d < (uint32_t)lut_size) w |= (uint32_t)l[d] << 24;
    return w;
  };
  const int64_t nvec = n >> 4;                       // (host: n % 16 == 0, 16-byte aligned planes)
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (int64_t)gridDim.x * blockDim.x) {
    int4 q[4];
    load_ids16(in + 16 * v, q);
    const uint4 gw = __ldcs(reinterpret_cast<const uint4*>(g) + v);
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) w[j] = look4(q[j]);
    __stcs(reinterpret_cast<uint4*>(out) + v, make_uint4(w[0], w[1], w[2], w[3]));
    h.word(w[0], gw.x); h.word(w[1], gw.y); h.word(w[2], gw.z); h.word(w[3], gw.w);
  }
  if (badid) atomicOr(status, LDIFF_STATUS_INST_RANGE);
  __syncthreads();
  h.flush(hist, C, status);
  if (PUSH) xchg_push_tail(C, px, s_step, gridDim.x * gridDim.y);
}

// One block per rank.  The j-th reduce of a rank adds the rows of every rank's j-th push: each thread
// owns a few (channel, bin) counters and, per source rank, spins until both words of the counter carry
// the step's tag.  A peer that never arrives (crashed rank) trips a timeout (10 s by default,
// ldiff_xchg_set_timeout) that sets LDIFF_STATUS_XCHG_TIMEOUT instead of hanging the GPU.
__global__ void __launch_bounds__(256)
xchg_reduce_kernel(XchgHeader* __restrict__ win, unsigned long long* __restrict__ out, int world,
                   int channels, int n, unsigned long long timeout_ns, int* __restrict__ status) {
  const unsigned long long want = win->reduced + 1;
  const unsigned long long tag = want & 0xffffffffull;
  const int slot = (int)(want % kXSlots);
  const unsigned long long t0 = globaltimer_ns();
  // a timeout that is already on record (crashed peer) is not waited for again: one timeout per failure
  bool late = (*reinterpret_cast<volatile int*>(status) & LDIFF_STATUS_XCHG_TIMEOUT) != 0;
  for (int i = threadIdx.x; i < channels * n; i += blockDim.x) {
    const int ch = i / n, bin = i - ch * n;
    unsigned long long s = 0;
    for (int src = 0; src < world; ++src) {
      const unsigned long long* w =
          xchg_row(reinterpret_cast<unsigned long long>(win), slot, src, ch, world, channels, n) + 2 * bin;
      unsigned long long w0 = ld_relaxed_sys(w), w1 = ld_relaxed_sys(w + 1);
      while (!late && ((w0 >> 32) != tag || (w1 >> 32) != tag)) {
        if (globaltimer_ns() - t0 > timeout_ns) { late = true; break; }
        __nanosleep(100);
        w0 = ld_relaxed_sys(w);
        w1 = ld_relaxed_sys(w + 1);
      }
      if ((w0 >> 32) == tag && (w1 >> 32) == tag) s += (w0 & 0xffffffffull) | (w1 << 32);
    }
    out[i] = s;
  }
  if (late) atomicOr(status, LDIFF_STATUS_XCHG_TIMEOUT);
  __syncthreads();                                   // every thread has read `reduced`
  if (threadIdx.x == 0) win->reduced = want;
}

// Stand-alone push: every channel's finished matrix (C[channels][n], e.g. the totals of a whole evaluation) goes to
// every rank's window in one launch of one block — the exchange without a producer kernel to ride on.
__global__ void __launch_bounds__(256)
xchg_push_kernel(const unsigned long long* __restrict__ C, XchgPush px) {
  for (int ch = 0; ch < px.channels; ++ch) {
    const unsigned long long step = px.win->step[ch] + 1;
    const unsigned long long tag = (step & 0xffffffffull) << 32;
    const int slot = (int)(step % kXSlots);
    for (int bin = threadIdx.x; bin < px.n; bin += blockDim.x) {
      const unsigned long long v = C[(int64_t)ch * px.n + bin];
      const unsigned long long w0 = (v & 0xffffffffull) | tag, w1 = (v >> 32) | tag;
      for (int q = 0; q < px.world; ++q) {
        unsigned long long* row = xchg_row(px.peers[q], slot, px.rank, ch, px.world, px.channels, px.n);
        st_relaxed_sys(row + 2 * bin, w0);
        st_relaxed_sys(row + 2 * bin + 1, w1);
      }
    }
    __syncthreads();                                   // every thread has read step[ch]
    if (threadIdx.x == 0) px.win->step[ch] = step;
  }
}

__global__ void __launch_bounds__(256)
labels_to_u8_kernel(const int64_t* __restrict__ in, uint8_t* __restrict__ out, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t nvec = n >> 2;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
    const longlong2 a = __ldcs(reinterpret_cast<const longlong2*>(in) + 2 * v);
    const longlong2 b = __ldcs(reinterpret_cast<const longlong2*>(in) + 2 * v + 1);
    const long long x[4] = {a.x, a.y, b.x, b.y};
    uint32_t w = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) w |= (uint32_t)((x[k] < 0 || x[k] > 254) ? 255 : x[k]) << (8 * k);
    reinterpret_cast<uint32_t*>(out)[v] = w;
  }
  const int64_t t = (nvec << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) out[t] = (uint8_t)((in[t] < 0 || in[t] > 254) ? 255 : in[t]);
}

}  // namespace ldiff

using namespace ldiff;

static int launch_confusion(const uint8_t* pred, const uint8_t* gt, const uint8_t* gt_lut, int64_t* C,
                            int64_t n_per_image, int n_images, int K, int* status, cudaStream_t st,
                            XchgPush px = XchgPush{}) {
  const int threads = 512;
  const int nbins = (K + 1) * K;
  int R = 32;
  while (R > 1 && (size_t)nbins * R * 4 > 64 * 1024) R >>= 1;
  const size_t smem = (size_t)nbins * R * 4;
  // blocks per image: 2 blocks of 512 threads per SM over the batch (measured best: fewer global
  // atomics per bin than 4, same streaming rate), each thread >= 1 vector
  const int64_t nvec = (n_per_image >> 4) > 0 ? (n_per_image >> 4) : 1;
  int64_t bx = (nvec + threads - 1) / threads;
  static const int bps = [] { const char* e = getenv("LDIFF_CONF_BPS"); return e ? atoi(e) : 2; }();
  const int64_t cap = ((int64_t)sm_count() * bps + n_images - 1) / n_images;
  if (bx > cap) bx = cap > 0 ? cap : 1;
  const dim3 grid((unsigned)bx, (unsigned)n_images);
  unsigned long long* Cu = reinterpret_cast<unsigned long long*>(C);
  const bool smallk = K <= 15;
#define CH3(RR, SK, LUT)                                                                            \
  do {                                                                                              \
    if (smem > 48 * 1024) {                                                                         \
      cudaFuncSetAttribute(confusion_hist_kernel<RR, SK, LUT>,                                      \
                           cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024);                 \
      cudaFuncSetAttribute(confusion_hist_push_kernel<RR, SK, LUT>,                                 \
                           cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024);                 \
    }                                                                                               \
    if (px.win)                                                                                     \
      confusion_hist_push_kernel<RR, SK, LUT><<<grid, threads, smem, st>>>(pred, gt, gt_lut, Cu,    \
                                                                           n_per_image, K, status,  \
                                                                           px);                     \
    else                                                                                            \
      confusion_hist_kernel<RR, SK, LUT><<<grid, threads, smem, st>>>(pred, gt, gt_lut, Cu,         \
                                                                      n_per_image, K, status);      \
  } while (0)
#define CH(RR, SK)                                \
  do {                                            \
    if (gt_lut) CH3(RR, SK, true);                \
    else CH3(RR, SK, false);                      \
  } while (0)
  if (smallk) {
    CH(32, true);                                  // K <= 15 -> nbins <= 240 -> R == 32 always
  } else {
    switch (R) {
      case 32: CH(32, false); break;
      case 16: CH(16, false); break;
      case 8: CH(8, false); break;
      case 4: CH(4, false); break;
      case 2: CH(2, false); break;
      default: CH(1, false); break;
    }
  }
#undef CH
#undef CH3
  return check_launch();
}

extern "C" int ldiff_confusion_hist(const uint8_t* pred, const uint8_t* gt, const uint8_t* gt_lut,
                                    int64_t* C, int64_t n, int K, int* status, void* stream) {
  if (!pred || !gt || !C || !status || n < 0 || K < 1) return LDIFF_EINVAL;
  if (K > 128) return LDIFF_EUNSUPPORTED;
  if (n == 0) return LDIFF_OK;
  return launch_confusion(pred, gt, gt_lut, C, n, 1, K, status, (cudaStream_t)stream);
}

extern "C" int ldiff_confusion_hist_batched(const uint8_t* pred, const uint8_t* gt,
                                            const uint8_t* gt_lut, int64_t* C, int64_t n_per_image,
                                            int n_images, int K, int* status, void* stream) {
  if (!pred || !gt || !C || !status || n_per_image < 0 || n_images < 0 || K < 1) return LDIFF_EINVAL;
  if (K > 128 || n_images > 65535) return LDIFF_EUNSUPPORTED;
  if (n_per_image == 0 || n_images == 0) return LDIFF_OK;
  return launch_confusion(pred, gt, gt_lut, C, n_per_image, n_images, K, status, (cudaStream_t)stream);
}

// ---- exchange window management (host side) ----------------------------------------------------
namespace {
struct Xchg {
  int world, rank, channels, n;
  size_t bytes;
  char* base;                          // this rank's window (cudaMalloc: IPC-exportable)
  void* mapped[kXMaxWorld];            // peers opened through CUDA IPC (closed in destroy)
  unsigned long long peers[kXMaxWorld];  // window base of every rank as mapped in this process
  unsigned long long timeout_ns;       // how long a reduce waits for a missing rank
};
}  // namespace

extern "C" int ldiff_xchg_create(int world, int rank, int channels, int n_i64, void** handle) {
  if (!handle || world < 1 || world > kXMaxWorld || rank < 0 || rank >= world || channels < 1 ||
      channels > kXMaxChan || n_i64 < 1)
    return LDIFF_EINVAL;
  Xchg* x = new Xchg();
  x->world = world; x->rank = rank; x->channels = channels; x->n = n_i64;
  x->timeout_ns = 10000000000ull;
  x->bytes = kXHeaderBytes + (size_t)kXSlots * world * channels * n_i64 * 2 * sizeof(int64_t);
  for (int i = 0; i < kXMaxWorld; ++i) { x->mapped[i] = nullptr; x->peers[i] = 0; }
  if (cudaMalloc(&x->base, x->bytes) != cudaSuccess || cudaMemset(x->base, 0, x->bytes) != cudaSuccess ||
      cudaDeviceSynchronize() != cudaSuccess) {
    cudaGetLastError();
    delete x;
    return LDIFF_ELAUNCH;
  }
  *handle = x;
  return LDIFF_OK;
}

extern "C" int ldiff_xchg_ipc_handle(void* handle, void* out64) {
  if (!handle || !out64) return LDIFF_EINVAL;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  if (cudaIpcGetMemHandle(&h, static_cast<Xchg*>(handle)->base) != cudaSuccess) {
    cudaGetLastError();
    return LDIFF_ELAUNCH;
  }
  memcpy(out64, &h, 64);
  return LDIFF_OK;
}

static int xchg_set_peers(Xchg* x, void* const* bases) {
  for (int q = 0; q < kXMaxWorld; ++q)
    x->peers[q] = q < x->world ? reinterpret_cast<unsigned long long>(bases[q]) : 0ull;
  return LDIFF_OK;
}

extern "C" int ldiff_xchg_connect_ipc(void* handle, const void* handles) {
  if (!handle || !handles) return LDIFF_EINVAL;
  Xchg* x = static_cast<Xchg*>(handle);
  void* bases[kXMaxWorld];
  for (int q = 0; q < x->world; ++q) {
    if (q == x->rank) { bases[q] = x->base; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const char*>(handles) + 64 * q, 64);
    if (cudaIpcOpenMemHandle(&x->mapped[q], h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      cudaGetLastError();
      return LDIFF_ELAUNCH;
    }
    bases[q] = x->mapped[q];
  }
  return xchg_set_peers(x, bases);
}

extern "C" int ldiff_xchg_connect_local(void* handle, void* const* peer_handles) {
  if (!handle || !peer_handles) return LDIFF_EINVAL;
  Xchg* x = static_cast<Xchg*>(handle);
  void* bases[kXMaxWorld];
  for (int q = 0; q < x->world; ++q) {
    if (!peer_handles[q]) return LDIFF_EINVAL;
    bases[q] = static_cast<Xchg*>(peer_handles[q])->base;
  }
  return xchg_set_peers(x, bases);
}

extern "C" int ldiff_xchg_set_timeout(void* handle, int64_t timeout_ms) {
  if (!handle || timeout_ms < 1) return LDIFF_EINVAL;
  static_cast<Xchg*>(handle)->timeout_ns = (unsigned long long)timeout_ms * 1000000ull;
  return LDIFF_OK;
}

extern "C" int ldiff_xchg_destroy(void* handle) {
  if (!handle) return LDIFF_EINVAL;
  Xchg* x = static_cast<Xchg*>(handle);
  cudaDeviceSynchronize();
  for (int q = 0; q < kXMaxWorld; ++q)
    if (x->mapped[q]) cudaIpcCloseMemHandle(x->mapped[q]);
  cudaFree(x->base);
  cudaGetLastError();
  delete x;
  return LDIFF_OK;
}

int ldiff::xchg_push_args(void* handle, int channel, int n_i64, XchgPush* px) {
  Xchg* x = static_cast<Xchg*>(handle);
  if (!x || channel < 0 || channel >= x->channels || x->n != n_i64) return LDIFF_EINVAL;
  if (!x->peers[x->rank]) return LDIFF_EINVAL;       // not connected yet
  *px = XchgPush{reinterpret_cast<XchgHeader*>(x->base), x->world, x->rank, x->channels, channel, x->n, {0}};
  for (int q = 0; q < x->world; ++q) px->peers[q] = x->peers[q];
  return LDIFF_OK;
}

template <typename IdT>
static int launch_lut_paint_hist(const IdT* inst, const uint8_t* lut, uint8_t* mask, const uint8_t* gt,
                                 int64_t* C, int64_t n_per_image, int B, int lut_size, int64_t lut_stride,
                                 int K, void* xchg, int channel, int* status, void* stream) {
  if (!inst || !lut || !mask || !gt || !C || !status || n_per_image < 0 || B < 0 || lut_size < 1 || K < 1)
    return LDIFF_EINVAL;
  if (K > 15 || B > 65535) return LDIFF_EUNSUPPORTED;  // byte-sized bins; larger K: ldiff_lut_paint + ldiff_confusion_hist
  XchgPush px{};
  if (xchg) {
    const int rc = xchg_push_args(xchg, channel, (K + 1) * K, &px);
    if (rc != LDIFF_OK) return rc;
  }
  if (n_per_image == 0 || B == 0) return xchg ? LDIFF_EINVAL : LDIFF_OK;   // a push needs a launch
  if (!aligned16(inst) || !aligned16(mask) || !aligned16(gt) || (n_per_image % 16)) return LDIFF_EALIGN;
  const int threads = 256;
  const size_t smem = (size_t)(K + 1) * K * 32 * 4;
  int64_t bx = ((n_per_image >> 4) + threads - 1) / threads;
  const int64_t cap = ((int64_t)sm_count() * 4 + B - 1) / B;     // 4 blocks of 256 threads per SM over the batch
  if (bx > cap) bx = cap > 0 ? cap : 1;
  const dim3 grid((unsigned)bx, (unsigned)B);
  unsigned long long* Cu = reinterpret_cast<unsigned long long*>(C);
  if (xchg)
    lut_paint_hist_kernel<true, IdT><<<grid, threads, smem, (cudaStream_t)stream>>>(inst, lut, mask, gt, Cu, n_per_image,
                                                                             lut_size, lut_stride, K, status, px);
  else
    lut_paint_hist_kernel<false, IdT><<<grid, threads, smem, (cudaStream_t)stream>>>(inst, lut, mask, gt, Cu, n_per_image,
                                                                              lut_size, lut_stride, K, status, px);
  return check_launch();
}

extern "C" int ldiff_lut_paint_hist(const int32_t* inst, const uint8_t* lut, uint8_t* mask, const uint8_t* gt,
                                    int64_t* C, int64_t n_per_image, int B, int lut_size, int64_t lut_stride,
                                    int K, void* xchg, int channel, int* status, void* stream) {
  return launch_lut_paint_hist(inst, lut, mask, gt, C, n_per_image, B, lut_size, lut_stride, K, xchg, channel, status,
                               stream);
}

extern "C" int ldiff_lut_paint_hist_u16(const uint16_t* inst, const uint8_t* lut, uint8_t* mask, const uint8_t* gt,
                                        int64_t* C, int64_t n_per_image, int B, int lut_size, int64_t lut_stride,
                                        int K, void* xchg, int channel, int* status, void* stream) {
  return launch_lut_paint_hist(inst, lut, mask, gt, C, n_per_image, B, lut_size, lut_stride, K, xchg, channel, status,
                               stream);
}

extern "C" int ldiff_confusion_hist_push(const uint8_t* pred, const uint8_t* gt, const uint8_t* gt_lut,
                                         int64_t* C, int64_t n, int K, void* xchg, int channel,
                                         int* status, void* stream) {
  if (!pred || !gt || !C || !status || !xchg || n < 1 || K < 1) return LDIFF_EINVAL;
  if (K > 128) return LDIFF_EUNSUPPORTED;
  XchgPush px{};
  const int rc = xchg_push_args(xchg, channel, (K + 1) * K, &px);
  if (rc != LDIFF_OK) return rc;
  return launch_confusion(pred, gt, gt_lut, C, n, 1, K, status, (cudaStream_t)stream, px);
}

extern "C" int ldiff_xchg_push(void* xchg, const int64_t* C, void* stream) {
  if (!xchg || !C) return LDIFF_EINVAL;
  Xchg* x = static_cast<Xchg*>(xchg);
  XchgPush px{};
  const int rc = xchg_push_args(xchg, 0, x->n, &px);
  if (rc != LDIFF_OK) return rc;
  xchg_push_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const unsigned long long*>(C), px);
  return check_launch();
}

extern "C" int ldiff_xchg_reduce(void* xchg, int64_t* out, int* status, void* stream) {
  if (!xchg || !out || !status) return LDIFF_EINVAL;
  Xchg* x = static_cast<Xchg*>(xchg);
  xchg_reduce_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<XchgHeader*>(x->base),
                                                          reinterpret_cast<unsigned long long*>(out), x->world,
                                                          x->channels, x->n, x->timeout_ns, status);
  return check_launch();
}

extern "C" int ldiff_labels_to_u8(const int64_t* in, uint8_t* out, int64_t n, void* stream) {
  if (!in || !out || n < 0) return LDIFF_EINVAL;
  if (n == 0) return LDIFF_OK;
  if (!aligned16(in) || (reinterpret_cast<uintptr_t>(out) & 3u)) return LDIFF_EALIGN;
  const int64_t items = (n >> 2) > (n & 3) ? (n >> 2) : (n & 3);
  labels_to_u8_kernel<<<grid_for(items, 256, 8), 256, 0, (cudaStream_t)stream>>>(in, out, n);
  return check_launch();
}
