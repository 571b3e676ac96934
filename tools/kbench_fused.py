"""Micro-benchmark of the round-2 fused kernels at the bench shape, next to the launches they replace."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.manual_seed(0)
from ldiffusion_b200 import _cabi, ops
from kbench import timeit

PEAK = 6555.5
dev = "cuda"
B, H, W, K = 8, 1024, 1024, 11
lib = _cabi.lib()
px = B * H * W


def line(name, us, byt):
    print(f"{name:44s} {us:8.2f} us  {byt / us / 1e3:8.1f} GB/s  {byt / us / 1e3 / PEAK:.3f} of peak", flush=True)


# lift + argmax (+ hist)
feat = torch.randn(B, 256, 32, 32, device=dev).bfloat16()
w = (torch.randn(K, 256, device=dev) / 16).bfloat16()
logits = ops.head_logits(feat, w, None)
smooth = torch.nn.functional.interpolate(torch.randn(B, K, 4, 4, device=dev) * 3, size=(32, 32), mode="bilinear").contiguous()
mask = torch.empty(B, H, W, dtype=torch.uint8, device=dev)
gt = torch.randint(0, K, (B, H, W), dtype=torch.uint8, device=dev)
C = torch.zeros(K + 1, K, dtype=torch.int64, device=dev)
for variant in (4, 1, 0):            # round-1 per-pixel kernel, envelope column form, envelope row form (default)
    lib.ldiff_tune(_cabi.TUNE_ARGMAX_VARIANT, variant)
    for nm, lg in (("random", logits), ("smooth", smooth)):
        line(f"lift_argmax variant={variant} {nm}", timeit(lambda i: ops._lift_argmax(lg, mask), 1), px)
for variant in (0,):
  lib.ldiff_tune(_cabi.TUNE_ARGMAX_VARIANT, variant)
  for nm, lg in (("random", logits), ("smooth", smooth)):
    line(f"lift_argmax_hist variant={variant} {nm}", timeit(lambda i: ops.lift_argmax_hist(lg, (H, W), gt, out=C, mask_out=mask), 1), 2 * px)
lib.ldiff_tune(_cabi.TUNE_ARGMAX_VARIANT, 0)
line("confusion_hist (stand-alone)", timeit(lambda i: ops.confusion_hist(mask.view(-1), gt.view(-1), K, out=C), 1), 2 * px)

# paint (+ hist)
insts = [torch.randint(0, 801, (B, H, W), dtype=torch.int32, device=dev) for _ in range(4)]
lut = torch.randint(0, K, (B, 801), dtype=torch.uint8, device=dev)
line("lut_paint", timeit(lambda i: ops.lut_paint(insts[i], lut, out=mask), 4), 5 * px)
line("lut_paint_hist", timeit(lambda i: ops.lut_paint_hist(insts[i], lut, gt, K, out=C, mask_out=mask), 4), 6 * px)
insts16 = [torch.from_numpy(t.cpu().numpy().astype("uint16")).to(dev) for t in insts]
line("lut_paint uint16 ids", timeit(lambda i: ops.lut_paint(insts16[i], lut, out=mask), 4), 3 * px)
line("lut_paint_hist uint16 ids", timeit(lambda i: ops.lut_paint_hist(insts16[i], lut, gt, K, out=C, mask_out=mask), 4), 4 * px)
del insts16

# decode tails
imgs = [torch.empty(B, 3, H, W, device=dev, dtype=torch.bfloat16).uniform_(-1.2, 1.2) for _ in range(6)]
planes = torch.empty(B, 6, H, W, dtype=torch.uint8, device=dev)
rgb = torch.empty(B, H, W, 3, dtype=torch.uint8, device=dev)
featc = torch.empty(B, 5, 64, 64, dtype=torch.bfloat16, device=dev)
lsmall = torch.empty(B, 1, 64, 64, dtype=torch.uint8, device=dev)
for tma in (0, 4, 6, 7, 8, 9, 10, 11, 12, 13):
    lib.ldiff_tune(_cabi.TUNE_DECODE_TAIL_TMA, tma)
    line(f"decode_tail_gray bf16 tma={tma}",
         timeit(lambda i: ops.decode_tail_gray(imgs[i], want_rgb=False, gray_out=planes[:, i % 5]), 6), px * 7)
    line(f"decode_tail_fused(+feat) bf16 tma={tma}",
         timeit(lambda i: ops.decode_tail_fused(imgs[i], planes[:, i % 5], feat_out=featc, feat_channel=i % 5), 6), px * 7)
    line(f"decode_tail_fused(+feat+rgb+label) bf16 tma={tma}",
         timeit(lambda i: ops.decode_tail_fused(imgs[i], planes[:, i % 5], rgb_out=rgb, feat_out=featc, feat_channel=i % 5,
                                                label=gt, label_plane_out=planes[:, 5], label_small_out=lsmall), 6), px * 12)
lib.ldiff_tune(_cabi.TUNE_DECODE_TAIL_TMA, 6)

# sampler: two launches vs one
n = B * 4 * 128 * 128
xs = [torch.randn(n, device=dev).bfloat16() for _ in range(8)]
o1, o2 = torch.empty_like(xs[0]), torch.empty_like(xs[0])
line("plms4 + laplace (2 launches) bf16 2MiB",
     timeit(lambda i: (ops.plms_step(xs[0], xs[1:5], 4, 1.0, -0.1, 0.5, out=o1), ops.laplace_qsample(xs[5], 0.5, seed=1, out=o2)), 1) , n * 2 * 8)
line("plms_step_noise mode 4 (1 launch) bf16 2MiB",
     timeit(lambda i: ops.plms_step_noise(xs[0], xs[1:5], 4, 1.0, -0.1, 0.5, xs[5], 0.5, seed=1, out=o1, noisy_out=o2), 1), n * 2 * 8)
ops.check_status(dev)

big = [torch.randn(1 << 27, device=dev).bfloat16() for _ in range(3)]
ob = torch.empty_like(big[0])
for r in (10, 7):
    lib.ldiff_tune(_cabi.TUNE_PHILOX_ROUNDS, r)
    line(f"laplace_qsample bf16 256 MiB, Philox4x32-{r}", timeit(lambda i: ops.laplace_qsample(big[i], 0.7, seed=1, offset=i, out=ob), 3, iters=20),
         2 * big[0].numel() * 2)
lib.ldiff_tune(_cabi.TUNE_PHILOX_ROUNDS, 0)
