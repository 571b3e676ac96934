// Widening N3: pixel-contrastive InfoNCE on per-pixel feature vectors (sm_100a).
//
// Reference: model/loss.py:44-109 (InfoNceLoss.compute_contrastive_loss).  For every sampled
// (anchor, positive, N negatives) triple the reference builds logits = [a.p, a.n_1 .. a.n_N] / T
// with a 1 x C times C x (1+N) matmul and adds cross_entropy(logits, target=0); the result is the
// mean over all triples.  That Python triple loop becomes one 128-thread block per triple: the anchor's C
// (= number of sampling steps, <= 16) features sit in registers, the 1+N candidates are gathered
// from the planar [C, h*w] feature map (logits stay in registers), a max / sum-exp log-softmax gives the loss, and
// the backward kernel recomputes the probabilities and scatters the gradient with fp32 atomics
// (candidate pixels repeat across triples).  Parity is with injected index sets; the sampling
// procedure itself (loss.py:64-87: 1 % of each class as anchors, one other pixel of the class as
// positive, N distinct pixels of other classes as negatives) is infonce_sample_kernel below — one
// block per image, counter-based randomness, no host round trip (the reference calls .item() and
// torch.randperm per anchor).
#include "common.cuh"

namespace ldiff {

constexpr int kMaxC = 16;

struct NceDims { int C; int64_t hw; int n_neg; int n_pairs; float inv_t; };

// feat: fp32 [B, C, hw]; pair p: batch b[p], anchor a[p], positive q[p], negatives neg[p, 0..n_neg)
__device__ __forceinline__ float nce_dot(const float* __restrict__ fb, const float (&a)[kMaxC], int C, int64_t hw,
                                         int idx) {
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < kMaxC; ++c)                    // full unroll + uniform guard: a[] stays in registers
    if (c < C) s = fmaf(a[c], __ldg(fb + (int64_t)c * hw + idx), s);
  return s;
}

constexpr int kNceThreads = 128;                     // one block per triple: 4 warps walk the 1+N candidates
constexpr int kNceCache = 9;                         // logits kept in registers per thread (covers N <= 1151)

__device__ __forceinline__ float nce_block_max(float v, float* s_red) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  v = fmaxf(fmaxf(s_red[0], s_red[1]), fmaxf(s_red[2], s_red[3]));
  __syncthreads();
  return v;
}
__device__ __forceinline__ float nce_block_sum(float v, float* s_red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  v = (s_red[0] + s_red[1]) + (s_red[2] + s_red[3]);
  __syncthreads();
  return v;
}

__global__ void __launch_bounds__(kNceThreads)
infonce_forward_kernel(const float* __restrict__ feat, const int* __restrict__ pb, const int* __restrict__ pa,
                       const int* __restrict__ pq, const int* __restrict__ neg, float* __restrict__ loss,
                       float* __restrict__ lse, NceDims d) {
  __shared__ float s_red[4];
  const int pair = blockIdx.x, tid = threadIdx.x;
  if (pb[pair] < 0) {                                              // unused slot of a padded pair list
    if (tid == 0) { loss[pair] = 0.f; lse[pair] = 0.f; }
    return;
  }
  const float* fb = feat + (int64_t)pb[pair] * d.C * d.hw;
  float a[kMaxC];
#pragma unroll
  for (int c = 0; c < kMaxC; ++c) a[c] = c < d.C ? __ldg(fb + (int64_t)c * d.hw + pa[pair]) : 0.f;
  const int* ng = neg + (int64_t)pair * d.n_neg;
  const int n = d.n_neg + 1;
  auto logit = [&](int j) { return nce_dot(fb, a, d.C, d.hw, j == 0 ? pq[pair] : ng[j - 1]) * d.inv_t; };
  float xs[kNceCache];
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < kNceCache; ++i) {
    const int j = tid + i * kNceThreads;
    xs[i] = j < n ? logit(j) : -INFINITY;
    m = fmaxf(m, xs[i]);
  }
  for (int j = tid + kNceCache * kNceThreads; j < n; j += kNceThreads) m = fmaxf(m, logit(j));
  m = nce_block_max(m, s_red);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kNceCache; ++i) s += expf(xs[i] - m);        // exp(-inf) = 0 for the padding
  for (int j = tid + kNceCache * kNceThreads; j < n; j += kNceThreads) s += expf(logit(j) - m);
  s = nce_block_sum(s, s_red);
  if (tid == 0) {
    const float l = m + logf(s);                                   // logsumexp of the logits
    lse[pair] = l;
    loss[pair] = l - xs[0];                                        // cross_entropy(logits, 0)
  }
}

// grad_feat += d(mean loss)/d(feat) * gscale, gscale = upstream gradient / n_pairs
__global__ void __launch_bounds__(kNceThreads)
infonce_backward_kernel(const float* __restrict__ feat, const int* __restrict__ pb, const int* __restrict__ pa,
                        const int* __restrict__ pq, const int* __restrict__ neg, const float* __restrict__ lse,
                        const float* __restrict__ gscale, float* __restrict__ grad, NceDims d) {
  __shared__ float s_ga[4][kMaxC];
  const int pair = blockIdx.x, tid = threadIdx.x;
  if (pb[pair] < 0) return;
  const int64_t boff = (int64_t)pb[pair] * d.C * d.hw;
  const float* fb = feat + boff;
  float* gb = grad + boff;
  const float g = __ldg(gscale) * d.inv_t;
  float a[kMaxC], ga[kMaxC];
#pragma unroll
  for (int c = 0; c < kMaxC; ++c) { a[c] = c < d.C ? __ldg(fb + (int64_t)c * d.hw + pa[pair]) : 0.f; ga[c] = 0.f; }
  const int* ng = neg + (int64_t)pair * d.n_neg;
  const int n = d.n_neg + 1;
  const float l = lse[pair];
  for (int j = tid; j < n; j += kNceThreads) {
    const int idx = j == 0 ? pq[pair] : ng[j - 1];
    float f[kMaxC], dot = 0.f;
#pragma unroll
    for (int c = 0; c < kMaxC; ++c) {                              // the candidate's features, loaded once
      f[c] = c < d.C ? __ldg(fb + (int64_t)c * d.hw + idx) : 0.f;
      dot = fmaf(a[c], f[c], dot);
    }
    const float w = (expf(dot * d.inv_t - l) - (j == 0 ? 1.f : 0.f)) * g;   // (softmax - onehot) * scale
#pragma unroll
    for (int c = 0; c < kMaxC; ++c) {
      ga[c] = fmaf(w, f[c], ga[c]);
      if (c < d.C) atomicAdd(gb + (int64_t)c * d.hw + idx, w * a[c]);
    }
  }
#pragma unroll
  for (int c = 0; c < kMaxC; ++c) {
    float v = ga[c];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((tid & 31) == 0) s_ga[tid >> 5][c] = v;
  }
  __syncthreads();
  if (tid < d.C)
    atomicAdd(gb + (int64_t)tid * d.hw + pa[pair], (s_ga[0][tid] + s_ga[1][tid]) + (s_ga[2][tid] + s_ga[3][tid]));
}

// ---- on-device sampling of the (anchor, positive, negatives) triples --------------------------------
constexpr int kNceLabels = 32;                       // label values must be < 32 (the reference's are 0..6)

__device__ __forceinline__ uint64_t nce_mix(uint64_t x) {          // splitmix64 finaliser
  x += 0x9e3779b97f4a7c15ull;
  x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
  return x ^ (x >> 31);
}
__device__ __forceinline__ uint64_t nce_key(uint64_t seed, uint64_t offset, int b, int label, int anchor, int what) {
  return nce_mix(nce_mix(seed ^ nce_mix(offset)) ^ ((uint64_t)b << 40) ^ ((uint64_t)label << 32) ^
                 ((uint64_t)(uint32_t)anchor << 4) ^ (uint64_t)what);
}
// keyed bijection of [0, m): 4-round Feistel network on the enclosing power of four, cycle-walked
// back into range.  x -> perm(x) for x = 0..k-1 is a draw of k DISTINCT values — the prefix of a
// random permutation, which is what torch.randperm(m)[:k] gives the reference.
__device__ __forceinline__ uint32_t nce_perm(uint32_t x, uint32_t m, uint64_t key) {
  int hb = 1;
  while ((1u << (2 * hb)) < m) ++hb;
  const uint32_t mask = (1u << hb) - 1u;
  do {
    uint32_t l = x >> hb, r = x & mask;
#pragma unroll
    for (int round = 0; round < 4; ++round) {
      const uint32_t f = (uint32_t)(nce_mix(key + ((uint64_t)r << 8) + (uint64_t)round) >> 17);
      const uint32_t t = l ^ (f & mask);
      l = r;
      r = t;
    }
    x = (l << hb) | r;
  } while (x >= m);
  return x;
}

// one block per image.  labels: uint8 [B, hw].  Pair slots of image b: [b*cap, (b+1)*cap); unused slots
// get pair_batch = -1.  n_valid[b] = triples of image b.
__global__ void __launch_bounds__(256)
infonce_sample_kernel(const uint8_t* __restrict__ labels, int hw, int n_neg, int cap, uint64_t seed,
                      uint64_t offset, int* __restrict__ pb, int* __restrict__ pa, int* __restrict__ pq,
                      int* __restrict__ neg, int* __restrict__ n_valid, int* __restrict__ status) {
  extern __shared__ uint16_t s_sorted[];             // pixel indices grouped by label, ascending inside a label
  __shared__ int s_cnt[kNceLabels], s_start[kNceLabels], s_aoff[kNceLabels + 1];
  __shared__ int s_wtot[8];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const uint8_t* lab = labels + (int64_t)b * hw;
  const int per = (hw + 255) / 256;                  // contiguous pixels per thread (keeps the grouping stable)
  const int p0 = min(hw, tid * per), p1 = min(hw, p0 + per);
  if (tid < kNceLabels) s_cnt[tid] = 0;
  __syncthreads();
  bool bad = false;
  for (int p = p0; p < p1; ++p) {
    const int l = lab[p];
    if (l >= kNceLabels) bad = true;
    else atomicAdd(&s_cnt[l], 1);
  }
  if (bad) atomicOr(status, LDIFF_STATUS_LABEL_RANGE);
  __syncthreads();
  if (tid == 0) {
    int st = 0, ao = 0;
    for (int l = 0; l < kNceLabels; ++l) {           // classes in ascending order, as torch.unique returns them
      s_start[l] = st;
      st += s_cnt[l];
      s_aoff[l] = ao;
      const bool valid = s_cnt[l] > 1 && hw - s_cnt[l] > n_neg;            // loss.py:73
      if (valid) ao = min(cap, ao + max(1, s_cnt[l] / 100));               // max(1, int(0.01 * len(pos_idx))), loss.py:75
    }
    s_aoff[kNceLabels] = ao;
    n_valid[b] = ao;
  }
  __syncthreads();
  for (int l = 0; l < kNceLabels; ++l) {             // stable grouping, one present label at a time
    if (s_cnt[l] == 0) continue;                     // (block-uniform)
    int mine = 0;
    for (int p = p0; p < p1; ++p) mine += lab[p] == l;
    int incl = mine;
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s_wtot[wid] = incl;
    __syncthreads();
    int base = s_start[l] + incl - mine;
    for (int w = 0; w < wid; ++w) base += s_wtot[w];
    for (int p = p0; p < p1; ++p)
      if (lab[p] == l) s_sorted[base++] = (uint16_t)p;
    __syncthreads();
  }
  const int total = s_aoff[kNceLabels];
  for (int pair = wid; pair < cap; pair += 8) {      // one warp per triple
    const int o = b * cap + pair;
    if (pair >= total) {                             // unused slot: fully defined, skipped downstream
      if (lane == 0) { pb[o] = -1; pa[o] = -1; pq[o] = -1; }
      for (int j = lane; j < n_neg; j += 32) neg[(int64_t)o * n_neg + j] = -1;
      continue;
    }
    int l = 0;
    while (s_aoff[l + 1] <= pair) ++l;               // the class this slot belongs to
    const int a = pair - s_aoff[l], cnt = s_cnt[l], st = s_start[l];
    const uint32_t arank = nce_perm((uint32_t)a, (uint32_t)cnt, nce_key(seed, offset, b, l, 0, 1));   // randperm(len)[:k]
    if (lane == 0) {
      uint32_t r = (uint32_t)(nce_key(seed, offset, b, l, a, 2) % (uint64_t)(cnt - 1));              // randint over the rest
      r += r >= arank;
      pb[o] = b;
      pa[o] = s_sorted[st + arank];
      pq[o] = s_sorted[st + r];
    }
    const uint32_t m = (uint32_t)(hw - cnt);
    const uint64_t kneg = nce_key(seed, offset, b, l, a, 3);
    int* ng = neg + (int64_t)o * n_neg;
    for (int j = lane; j < n_neg; j += 32) {
      const uint32_t rj = nce_perm((uint32_t)j, m, kneg);                   // randperm(len(neg_idx))[:N]
      ng[j] = s_sorted[rj < (uint32_t)st ? rj : rj + cnt];                 // every pixel outside the class's segment
    }
  }
}

}  // namespace ldiff

using namespace ldiff;

extern "C" int ldiff_infonce_sample(const uint8_t* labels, int B, int64_t hw, int n_neg, int cap, uint64_t seed,
                                    uint64_t offset, int* pair_batch, int* pair_anchor, int* pair_pos,
                                    int* pair_neg, int* n_valid, int* status, void* stream) {
  if (!labels || !pair_batch || !pair_anchor || !pair_pos || !pair_neg || !n_valid || !status || B < 0 || hw < 2 ||
      n_neg < 1 || cap < 1)
    return LDIFF_EINVAL;
  if (hw > 16384) return LDIFF_EUNSUPPORTED;         // uint16 indices in 32 KB of shared memory
  if (B == 0) return LDIFF_OK;
  infonce_sample_kernel<<<B, 256, (size_t)hw * sizeof(uint16_t), (cudaStream_t)stream>>>(
      labels, (int)hw, n_neg, cap, seed, offset, pair_batch, pair_anchor, pair_pos, pair_neg, n_valid, status);
  return check_launch();
}

static int nce_check(const void* feat, const int* pb, const int* pa, const int* pq, const int* neg, int C, int64_t hw,
                     int n_neg, int n_pairs, float temperature) {
  if (!feat || !pb || !pa || !pq || !neg || C < 1 || hw < 1 || n_neg < 0 || n_pairs < 0 || !(temperature > 0.f))
    return LDIFF_EINVAL;
  if (C > kMaxC) return LDIFF_EUNSUPPORTED;
  return LDIFF_OK;
}

extern "C" int ldiff_infonce_forward(const float* feat, const int* pair_batch, const int* pair_anchor,
                                     const int* pair_pos, const int* pair_neg, float* loss_per_pair,
                                     float* lse_per_pair, int C, int64_t hw, int n_neg, int n_pairs,
                                     float temperature, void* stream) {
  const int rc = nce_check(feat, pair_batch, pair_anchor, pair_pos, pair_neg, C, hw, n_neg, n_pairs, temperature);
  if (rc != LDIFF_OK || !loss_per_pair || !lse_per_pair) return rc != LDIFF_OK ? rc : LDIFF_EINVAL;
  if (n_pairs == 0) return LDIFF_OK;
  NceDims d{C, hw, n_neg, n_pairs, 1.f / temperature};
  infonce_forward_kernel<<<n_pairs, kNceThreads, 0, (cudaStream_t)stream>>>(
      feat, pair_batch, pair_anchor, pair_pos, pair_neg, loss_per_pair, lse_per_pair, d);
  return check_launch();
}

extern "C" int ldiff_infonce_backward(const float* feat, const int* pair_batch, const int* pair_anchor,
                                      const int* pair_pos, const int* pair_neg, const float* lse_per_pair,
                                      const float* grad_scale, float* grad_feat, int C, int64_t hw, int n_neg,
                                      int n_pairs, float temperature, void* stream) {
  const int rc = nce_check(feat, pair_batch, pair_anchor, pair_pos, pair_neg, C, hw, n_neg, n_pairs, temperature);
  if (rc != LDIFF_OK || !lse_per_pair || !grad_scale || !grad_feat) return rc != LDIFF_OK ? rc : LDIFF_EINVAL;
  if (n_pairs == 0) return LDIFF_OK;
  NceDims d{C, hw, n_neg, n_pairs, 1.f / temperature};
  infonce_backward_kernel<<<n_pairs, kNceThreads, 0, (cudaStream_t)stream>>>(
      feat, pair_batch, pair_anchor, pair_pos, pair_neg, lse_per_pair, grad_scale, grad_feat, d);
  return check_launch();
}
