// Adjoint of the a-4 bilinear lift (+ weighted gray), for the training caller of the path
// (ldiffusion.py:240-252: the InfoNCE loss on the lifted gray features back-propagates into the
// decoder output).  Round-2 widening, not measured yet; only reached through autograd.
//
//   forward : out[b, ch, oy, ox] = sum_c w_c * sum_{ty, tx} ly * lx * src[b, c, y, x]      (w_c = 1 without gray)
//   adjoint : grad_src[b, c, y, x] += w_c * ly * lx * grad_out[b, ch, oy, ox]
//
// One thread per output pixel scatters its <= 4 taps per channel with fp32 atomics into a
// caller-zeroed grad_src.  Down-sampling by an integer factor >= 2 (the path's 1024 -> 64) has
// disjoint footprints, so the sums are order-independent there; elsewhere the atomics make the
// last bits run-dependent.  HBM traffic is the memset of grad_src plus 4/256 of it again.
#include "common.cuh"

namespace ldiff {
namespace bwd {

struct Axis { float scale; int in, out; };
struct Tap { int i0, i1; float l0, l1; };

__device__ __forceinline__ Tap make_tap(const Axis& a, int dst) {   // as bilinear.cu::make_tap
  Tap t;
  if (a.in == a.out) { t.i0 = t.i1 = dst; t.l0 = 1.f; t.l1 = 0.f; return t; }
  float src = fmaxf(__fmaf_rn(a.scale, (float)dst + 0.5f, -0.5f), 0.f);
  t.i0 = min((int)floorf(src), a.in - 1);
  t.i1 = min(t.i0 + 1, a.in - 1);
  t.l1 = fminf(fmaxf(__fsub_rn(src, (float)t.i0), 0.f), 1.f);
  t.l0 = __fsub_rn(1.f, t.l1);
  return t;
}

// grad_out: [B, Ctot, out_h, out_w] fp32, channels [dch, dch + (gray ? 1 : C)) are this lift's;
// grad_src: [B, C, in_h, in_w] fp32 with element strides (sbs, scs), accumulated into.
__global__ void __launch_bounds__(256)
lift_backward_kernel(const float* __restrict__ grad_out, int Ctot, int dch, float* __restrict__ grad_src,
                     int C, int64_t sbs, int64_t scs, Axis ay, Axis ax, int B, int gray) {
  const int64_t HW = (int64_t)ay.out * ax.out;
  const int64_t total = HW * B;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const float wgray[3] = {0.2989f, 0.5870f, 0.1140f};
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int b = (int)(i / HW);
    const int64_t p = i - (int64_t)b * HW;
    const int y = (int)(p / ax.out), x = (int)(p - (int64_t)y * ax.out);
    const Tap ty = make_tap(ay, y), tx = make_tap(ax, x);
    const float w00 = ty.l0 * tx.l0, w01 = ty.l0 * tx.l1, w10 = ty.l1 * tx.l0, w11 = ty.l1 * tx.l1;
    const float g_gray = gray ? __ldg(grad_out + ((int64_t)b * Ctot + dch) * HW + p) : 0.f;
    for (int c = 0; c < C; ++c) {
      const float g = gray ? wgray[c] * g_gray : __ldg(grad_out + ((int64_t)b * Ctot + dch + c) * HW + p);
      float* sc = grad_src + b * sbs + c * scs;
      float* r0 = sc + (int64_t)ty.i0 * ax.in;
      float* r1 = sc + (int64_t)ty.i1 * ax.in;
      atomicAdd(r0 + tx.i0, w00 * g);
      if (w01 != 0.f) atomicAdd(r0 + tx.i1, w01 * g);
      if (w10 != 0.f) atomicAdd(r1 + tx.i0, w10 * g);
      if (w11 != 0.f) atomicAdd(r1 + tx.i1, w11 * g);
    }
  }
}

}  // namespace bwd
}  // namespace ldiff

using namespace ldiff;

extern "C" int ldiff_bilinear_lift_backward(const float* grad_out, int Ctot, int dst_channel, int H, int W,
                                            float* grad_src, int C, int h, int w, int64_t src_batch_stride,
                                            int64_t src_channel_stride, int B, int gray, void* stream) {
  if (!grad_out || !grad_src || Ctot < 1 || dst_channel < 0 || H < 1 || W < 1 || C < 1 || h < 1 || w < 1 || B < 0)
    return LDIFF_EINVAL;
  if (gray && C != 3) return LDIFF_EINVAL;
  if (dst_channel + (gray ? 1 : C) > Ctot) return LDIFF_EINVAL;
  if (src_channel_stride < (int64_t)h * w || src_batch_stride < src_channel_stride * C) return LDIFF_EINVAL;
  if (B == 0) return LDIFF_OK;
  bwd::Axis ay{(float)h / (float)H, h, H}, ax{(float)w / (float)W, w, W};
  const int64_t total = (int64_t)H * W * B;
  bwd::lift_backward_kernel<<<grid_for(total, 256, 8), 256, 0, (cudaStream_t)stream>>>(
      grad_out, Ctot, dst_channel, grad_src, C, src_batch_stride, src_channel_stride, ay, ax, B, gray);
  return check_launch();
}
