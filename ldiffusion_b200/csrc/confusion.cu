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

#include "common.cuh"

namespace ldiff {

// bytes of w that are >= k (k <= 128) get 0x80, others 0
__device__ __forceinline__ uint32_t bytes_ge(uint32_t w, uint32_t k) {
  return (((w & 0x7f7f7f7fu) + (0x80u - k) * 0x01010101u) | w) & 0x80808080u;
}

// ---- peer exchange window (multi-GPU): one per rank, mapped into every peer over NVLink -------------
// The only cross-rank step of the path is the SUM of the int64 matrices.  Instead of a separate
// collective, the LAST block of the histogram kernel stores the finished matrix straight into every
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
  unsigned int ticket[kXMaxChan];                              // last-block election of the histogram grid
};
static_assert(sizeof(XchgHeader) <= kXHeaderBytes, "header does not fit");
struct XchgPush {                                              // by-value kernel argument; win == nullptr: off
  XchgHeader* win;
  int world, rank, channels, channel, n;
  unsigned long long peers[kXMaxWorld];                        // window base of every rank, as mapped here
};
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
  for (int i = threadIdx.x; i < nbins * R; i += blockDim.x) hist[i] = 0;
  if (HAS_LUT)
    for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = gt_lut[i];
  // (push) this launch's step number: only the previous launch's last block ever changes the counter,
  // so it is read here, long before the tail needs it, instead of on the tail's dependent chain
  __shared__ unsigned long long s_step;
  if (PUSH && threadIdx.x == 0) s_step = __ldcg(&px.win->step[px.channel]) + 1;
  __syncthreads();

  // grid.y = image (batched form: one matrix per image)
  pred += (int64_t)blockIdx.y * n;
  gt += (int64_t)blockIdx.y * n;
  C += (int64_t)blockIdx.y * nbins;

  uint32_t* my = hist + (threadIdx.x & (R - 1));     // this lane's replica column
  uint32_t bad = 0;
  const uint32_t k4 = (uint32_t)K * 0x01010101u;

  auto pixel = [&](uint32_t p, uint32_t g) {         // generic path
    if (HAS_LUT) g = lut[g];
    g = min(g, (uint32_t)K);
    if (p >= (uint32_t)K) { bad = 1; p = 0; }
    atomicAdd(my + (g * K + p) * R, 1u);
  };
  auto word = [&](uint32_t pw, uint32_t gw) {
    if (SMALLK) {
      if (HAS_LUT)
        gw = (uint32_t)lut[gw & 0xff] | ((uint32_t)lut[(gw >> 8) & 0xff] << 8) |
             ((uint32_t)lut[(gw >> 16) & 0xff] << 16) | ((uint32_t)lut[gw >> 24] << 24);
      const uint32_t gm = (bytes_ge(gw, K) >> 7) * 0xffu;          // 0xff where gt >= K
      const uint32_t gc = (gw & ~gm) | (k4 & gm);
      const uint32_t po = bytes_ge(pw, K);
      bad |= po;
      const uint32_t pc = pw & ~((po >> 7) * 0xffu);
      const uint32_t bins = gc * (uint32_t)K + pc;                  // four byte-sized bin indices
      atomicAdd(my + (bins & 0xffu) * R, 1u);
      atomicAdd(my + ((bins >> 8) & 0xffu) * R, 1u);
      atomicAdd(my + ((bins >> 16) & 0xffu) * R, 1u);
      atomicAdd(my + (bins >> 24) * R, 1u);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) pixel((pw >> (8 * k)) & 0xffu, (gw >> (8 * k)) & 0xffu);
    }
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
  if (bad) atomicOr(status, LDIFF_STATUS_PRED_RANGE);
  __syncthreads();

  for (int bin = threadIdx.x; bin < nbins; bin += blockDim.x) {
    unsigned long long s = 0;
#pragma unroll 8
    for (int r = 0; r < R; ++r) s += hist[bin * R + ((r + threadIdx.x) & (R - 1))];
    if (s) atomicAdd(C + bin, s);
  }
  if (!PUSH) return;

  // ---- fused exchange: the last block to finish owns the complete matrix and pushes it to every rank
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0)
    s_last = atomicAdd(&px.win->ticket[px.channel], 1u) == gridDim.x * gridDim.y - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const unsigned long long step = s_step;
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
    px.win->ticket[px.channel] = 0;                  // re-armed for the next launch (stream-ordered)
  }
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

extern "C" int ldiff_confusion_hist_push(const uint8_t* pred, const uint8_t* gt, const uint8_t* gt_lut,
                                         int64_t* C, int64_t n, int K, void* xchg, int channel,
                                         int* status, void* stream) {
  if (!pred || !gt || !C || !status || !xchg || n < 1 || K < 1) return LDIFF_EINVAL;
  if (K > 128) return LDIFF_EUNSUPPORTED;
  Xchg* x = static_cast<Xchg*>(xchg);
  if (channel < 0 || channel >= x->channels || x->n != (K + 1) * K) return LDIFF_EINVAL;
  if (!x->peers[x->rank]) return LDIFF_EINVAL;       // not connected yet
  XchgPush px{reinterpret_cast<XchgHeader*>(x->base), x->world, x->rank, x->channels, channel, x->n, {0}};
  for (int q = 0; q < x->world; ++q) px.peers[q] = x->peers[q];
  return launch_confusion(pred, gt, gt_lut, C, n, 1, K, status, (cudaStream_t)stream, px);
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
