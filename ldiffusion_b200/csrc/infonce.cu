// Widening N3: pixel-contrastive InfoNCE on per-pixel feature vectors (sm_100a).
//
// Reference: model/loss.py:44-109 (InfoNceLoss.compute_contrastive_loss).  For every sampled
// (anchor, positive, N negatives) triple the reference builds logits = [a.p, a.n_1 .. a.n_N] / T
// with a 1 x C times C x (1+N) matmul and adds cross_entropy(logits, target=0); the result is the
// mean over all triples.  That Python triple loop becomes one warp per triple: the anchor's C
// (= number of sampling steps, <= 16) features sit in registers, the 1+N candidates are gathered
// from the planar [C, h*w] feature map, a two-pass (max, sum-exp) log-softmax gives the loss, and
// the backward kernel recomputes the probabilities and scatters the gradient with fp32 atomics
// (candidate pixels repeat across triples).  Sampling stays on the host (torch.randperm in the
// reference); parity is with injected index sets.
#include "common.cuh"

namespace ldiff {

constexpr int kMaxC = 16;

struct NceDims { int C; int64_t hw; int n_neg; int n_pairs; float inv_t; };

// feat: fp32 [B, C, hw]; pair p: batch b[p], anchor a[p], positive q[p], negatives neg[p, 0..n_neg)
__device__ __forceinline__ float nce_dot(const float* __restrict__ fb, const float (&a)[kMaxC], int C, int64_t hw,
                                         int idx) {
  float s = 0.f;
  for (int c = 0; c < C; ++c) s = fmaf(a[c], __ldg(fb + (int64_t)c * hw + idx), s);
  return s;
}

__global__ void __launch_bounds__(256)
infonce_forward_kernel(const float* __restrict__ feat, const int* __restrict__ pb, const int* __restrict__ pa,
                       const int* __restrict__ pq, const int* __restrict__ neg, float* __restrict__ loss,
                       float* __restrict__ lse, NceDims d) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= d.n_pairs) return;
  const float* fb = feat + (int64_t)pb[warp] * d.C * d.hw;
  float a[kMaxC];
  for (int c = 0; c < d.C; ++c) a[c] = __ldg(fb + (int64_t)c * d.hw + pa[warp]);
  const int* ng = neg + (int64_t)warp * d.n_neg;
  const int n = d.n_neg + 1;
  float m = -INFINITY;
  for (int j = lane; j < n; j += 32) {
    const int idx = j == 0 ? pq[warp] : ng[j - 1];
    m = fmaxf(m, nce_dot(fb, a, d.C, d.hw, idx) * d.inv_t);
  }
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float s = 0.f;
  for (int j = lane; j < n; j += 32) {
    const int idx = j == 0 ? pq[warp] : ng[j - 1];
    s += expf(nce_dot(fb, a, d.C, d.hw, idx) * d.inv_t - m);
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) {
    const float l = m + logf(s);                                   // logsumexp of the logits
    lse[warp] = l;
    loss[warp] = l - nce_dot(fb, a, d.C, d.hw, pq[warp]) * d.inv_t;  // cross_entropy(logits, 0)
  }
}

// grad_feat += d(mean loss)/d(feat) * gscale, gscale = upstream gradient / n_pairs
__global__ void __launch_bounds__(256)
infonce_backward_kernel(const float* __restrict__ feat, const int* __restrict__ pb, const int* __restrict__ pa,
                        const int* __restrict__ pq, const int* __restrict__ neg, const float* __restrict__ lse,
                        const float* __restrict__ gscale, float* __restrict__ grad, NceDims d) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= d.n_pairs) return;
  const int64_t boff = (int64_t)pb[warp] * d.C * d.hw;
  const float* fb = feat + boff;
  float* gb = grad + boff;
  const float g = __ldg(gscale) * d.inv_t;
  float a[kMaxC], ga[kMaxC];
  for (int c = 0; c < d.C; ++c) { a[c] = __ldg(fb + (int64_t)c * d.hw + pa[warp]); ga[c] = 0.f; }
  const int* ng = neg + (int64_t)warp * d.n_neg;
  const int n = d.n_neg + 1;
  const float l = lse[warp];
  for (int j = lane; j < n; j += 32) {
    const int idx = j == 0 ? pq[warp] : ng[j - 1];
    const float p = expf(nce_dot(fb, a, d.C, d.hw, idx) * d.inv_t - l) - (j == 0 ? 1.f : 0.f);   // softmax - onehot
    const float w = p * g;
    for (int c = 0; c < d.C; ++c) {
      ga[c] = fmaf(w, __ldg(fb + (int64_t)c * d.hw + idx), ga[c]);
      atomicAdd(gb + (int64_t)c * d.hw + idx, w * a[c]);
    }
  }
  for (int c = 0; c < d.C; ++c) {
    float v = ga[c];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) atomicAdd(gb + (int64_t)c * d.hw + pa[warp], v);
  }
}

}  // namespace ldiff

using namespace ldiff;

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
  infonce_forward_kernel<<<(n_pairs * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
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
  infonce_backward_kernel<<<(n_pairs * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      feat, pair_batch, pair_anchor, pair_pos, pair_neg, lse_per_pair, grad_scale, grad_feat, d);
  return check_launch();
}
