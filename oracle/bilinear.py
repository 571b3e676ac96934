"""Oracle for stage a-4: bilinear lift + gray + per-step concat.

Test infrastructure only (see ``oracle/__init__.py``).

Reference: ``ldiffusion.py:240-251`` (decode -> bilinear 64x64 -> weighted gray
-> ``torch.cat`` over steps; last RGB -> bilinear 1024x1024) and the label path
``ldiffusion.py:224-226`` (uint8 -> float -> bilinear 64x64 -> uint8 truncation).
Same primitive at ``conductor.py:109,135,281,285,293``, ``loss.py:35-36``,
``segmentor.py:324-326,340``.  ``F.interpolate(mode='bilinear',
align_corners=False)`` is torch (third-party, importable): the chain tier calls
it; the spec tier restates ATen's ``upsample_bilinear2d`` arithmetic with every
fp32 rounding pinned (no FMA contraction), which is what the CUDA kernel does.
"""
import numpy as np
import torch
import torch.nn.functional as F

from ._fp import add32, fma32, mul32, sub32

GRAY_W = (0.2989, 0.5870, 0.1140)      # ldiffusion.py:241


# ---------------------------- chain tier ----------------------------------

def lift_chain(x, size):
    return F.interpolate(x, size=size, mode="bilinear", align_corners=False)


def gray_weighted_chain(rgb):
    """ldiffusion.py:241-242."""
    w = torch.tensor(GRAY_W, dtype=torch.float32).view(1, 3, 1, 1)
    return (rgb * w).sum(dim=1, keepdim=True)


def feature_concat_chain(decoded_steps, size=(64, 64)):
    """ldiffusion.py:240-247: per step lift -> gray -> cat on dim 1."""
    out = None
    for d in decoded_steps:
        g = gray_weighted_chain(lift_chain(d.to(torch.float32), size).to(torch.float32))
        out = g if out is None else torch.cat([out, g], dim=1)
    return out


def label_down_chain(label_u8, size=(64, 64)):
    """ldiffusion.py:224-226."""
    return lift_chain(label_u8.to(torch.float32), size).to(torch.uint8)


# ---------------------------- spec tier ------------------------------------
# Pinned to what ATen's CPU kernel computes in this image (torch 2.11, x86,
# GCC with FMA contraction), found by exhaustive comparison of contraction
# patterns (tests/test_oracle_bilinear.py keeps the evidence):
#   src   = fma(scale, dst + 0.5, -0.5)            scale = fl(in / out)
#   t     = fma(w0, a, fl(w1 * b))                 horizontal, per source row
#   v     = fma(h0, t_top, fl(h1 * t_bottom))      vertical
# This reproduces F.interpolate bit for bit on the up-sampling shapes of the
# path (32->1024 logits, 64->1024 RGB, and e.g. 37x53 -> 101x77).  ATen has no
# single arithmetic: which contraction its compiler chose depends on the loop
# specialisation a shape lands in (its 16x down-sampling uses a 4-term sum, a
# 13x9 -> 31x40 lift a mirrored FMA); there the spec agrees with it to <= 2 ulp
# and features are compared within the 1e-3 relative contract.

def _f32(x):
    return np.asarray(x, dtype=np.float32)


def source_index(out_size, in_size):
    """ATen area_pixel_compute_source_index (align_corners=False, not cubic) +
    guard_index_and_lambda, fp32.  Returns (i0, i1, lam0, lam1)."""
    if out_size == in_size:                       # ATen copies when scale == 1
        i = np.arange(out_size, dtype=np.int64)
        return i, i, np.ones(out_size, np.float32), np.zeros(out_size, np.float32)
    scale = np.float32(in_size) / np.float32(out_size)
    dst = np.arange(out_size, dtype=np.float32)
    src = fma32(np.full(out_size, scale, np.float32), dst + np.float32(0.5),
                np.full(out_size, -0.5, np.float32))
    src = np.maximum(src, np.float32(0.0))
    i0 = np.minimum(np.floor(src).astype(np.int64), in_size - 1)
    i1 = np.minimum(i0 + 1, in_size - 1)
    lam1 = np.minimum(np.maximum(sub32(src, i0.astype(np.float32)), np.float32(0)), np.float32(1))
    lam0 = sub32(np.float32(1.0), lam1)
    return i0, i1, lam0, lam1


def lift_spec(x, size):
    """x: float array [..., h, w] -> fp32 [..., H, W]."""
    x = _f32(x)
    H, W = size
    h, w = x.shape[-2:]
    y0, y1, hy0, hy1 = source_index(H, h)
    x0, x1, wx0, wx1 = source_index(W, w)
    rows = x[..., :, x0], x[..., :, x1]
    t = fma32(wx0, rows[0], mul32(wx1, rows[1]))           # [..., h, W] horizontal pass
    top = t[..., y0, :]
    bot = t[..., y1, :]
    return fma32(hy0[:, None], top, mul32(hy1[:, None], bot))


def gray_weighted_spec(rgb):
    """[B,3,H,W] -> [B,1,H,W]: fl(fl(fl(wr*R)+fl(wg*G))+fl(wb*B)) (no FMA: the
    chain is three separate eager kernels, mul then a sequential sum)."""
    rgb = _f32(rgb)
    wr, wg, wb = (np.float32(v) for v in GRAY_W)
    s = add32(mul32(wr, rgb[:, 0]), mul32(wg, rgb[:, 1]))
    return add32(s, mul32(wb, rgb[:, 2]))[:, None]


def feature_concat_spec(decoded_steps, size=(64, 64)):
    return np.concatenate([gray_weighted_spec(lift_spec(_f32(d), size)) for d in decoded_steps], axis=1)


def label_down_spec(label_u8, size=(64, 64)):
    """float -> uint8 conversion truncates toward zero (values are in [0,255])."""
    return lift_spec(np.asarray(label_u8).astype(np.float32), size).astype(np.uint8)


def feature_concat_grad_chain(decoded_steps, grad_out, size=(64, 64)):
    """torch-CPU autograd through the literal chain of ldiffusion.py:240-247 (interpolate -> weighted gray ->
    cat): gradients w.r.t. every decoded step for an upstream ``grad_out`` [B,n,h,w]."""
    import torch
    import torch.nn.functional as F
    xs = [d.detach().clone().float().requires_grad_(True) for d in decoded_steps]
    chans = []
    for x in xs:
        r = F.interpolate(x, size=size, mode="bilinear", align_corners=False)
        chans.append((0.2989 * r[:, 0] + 0.5870 * r[:, 1] + 0.1140 * r[:, 2]).unsqueeze(1))
    torch.cat(chans, dim=1).backward(grad_out)
    return [x.grad for x in xs]


def lift_backward_spec(grad_out, src_shape, gray=False):
    """What ``ldiff_bilinear_lift_backward`` scatters, restated with numpy: the forward's taps
    (``source_index``), weight ``ly * lx`` per tap, ``GRAY_W[c]`` per channel in gray mode, fp32 products,
    accumulation in output-pixel order (the kernel's atomics may add in another order where footprints
    overlap).  grad_out: [B, (1 if gray else C), H, W] -> [B, C, h, w]."""
    g = _f32(grad_out)
    B, C, h, w = src_shape
    H, W = g.shape[2:]
    yi0, yi1, yl0, yl1 = source_index(H, h)
    xi0, xi1, xl0, xl1 = source_index(W, w)
    out = np.zeros((B, C, h, w), np.float32)
    wy = [(yi0, yl0), (yi1, yl1)]
    wx = [(xi0, xl0), (xi1, xl1)]
    for c in range(C):
        gc = g[:, 0] * np.float32(GRAY_W[c]) if gray else g[:, c]
        for iy, ly in wy:
            for ix, lx in wx:
                wgt = (ly[:, None] * lx[None, :]).astype(np.float32)          # [H, W]
                contrib = (wgt[None] * gc).astype(np.float32)                  # [B, H, W]
                np.add.at(out[:, c], (slice(None), iy[:, None], ix[None, :]), contrib)
    return out
