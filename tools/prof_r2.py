"""Launch the round-2 kernels a few times at the bench shape (for ncu -k regex:...)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ldiffusion_b200 import ops

dev = "cuda"
B, H, W, K = 8, 1024, 1024, 11
feat = torch.randn(B, 256, 32, 32, device=dev).bfloat16()
w = (torch.randn(K, 256, device=dev) / 16).bfloat16()
mask = torch.empty(B, H, W, dtype=torch.uint8, device=dev)
gt = torch.randint(0, K, (B, H, W), dtype=torch.uint8, device=dev)
C = torch.zeros(K + 1, K, dtype=torch.int64, device=dev)
inst = torch.randint(0, 801, (B, H, W), dtype=torch.int32, device=dev)
lut = torch.randint(0, K, (B, 801), dtype=torch.uint8, device=dev)
imgs = [torch.empty(B, 3, H, W, device=dev, dtype=torch.bfloat16).uniform_(-1.2, 1.2) for _ in range(2)]
planes = torch.empty(B, 6, H, W, dtype=torch.uint8, device=dev)
rgb = torch.empty(B, H, W, 3, dtype=torch.uint8, device=dev)
featc = torch.empty(B, 5, 64, 64, dtype=torch.bfloat16, device=dev)
lsmall = torch.empty(B, 1, 64, 64, dtype=torch.uint8, device=dev)
small = torch.empty(B, 3, 64, 64, dtype=torch.bfloat16, device=dev)
up = torch.empty(B, 3, H, W, dtype=torch.bfloat16, device=dev)
n = B * 4 * 128 * 128
xs = [torch.randn(n, device=dev).bfloat16() for _ in range(6)]
o1, o2 = torch.empty_like(xs[0]), torch.empty_like(xs[0])
inst_feats = torch.randn(B, 800, 256, device=dev).bfloat16()
ids = torch.arange(1, 801, dtype=torch.int32, device=dev)
big = torch.randn(1 << 26, device=dev).bfloat16()
bigo = torch.empty_like(big)
# the decode tails in both shapes the pass uses: 2 stages x 2 CTAs/SM (a single pass) and 2 x 1 (the ring of passes)
from ldiffusion_b200 import _cabi
lib = _cabi.lib()
for it in range(3):
    if it == 2:                                              # ncu --profile-from-start off: only the warm iteration
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
    logits = ops.head_logits(feat, w, None)
    ops._lift_argmax(logits, mask)
    ops.confusion_hist(mask.view(-1), gt.view(-1), K, out=C)
    ops.lift_argmax_hist(logits, (H, W), gt, out=C, mask_out=mask)
    ops.cell_classify(inst_feats, w, None, ids, 801)
    ops.lut_paint_hist(inst, lut, gt, K, out=C, mask_out=mask)
    ops.laplace_qsample(big, 0.7, seed=1, offset=it, out=bigo)
    for shape in (6, 11):
        lib.ldiff_tune(_cabi.TUNE_DECODE_TAIL_TMA, shape)
        ops.decode_tail_fused(imgs[it % 2], planes[:, it], feat_out=featc, feat_channel=it)
    lib.ldiff_tune(_cabi.TUNE_DECODE_TAIL_TMA, 6)
    ops.decode_tail_fused(imgs[it % 2], planes[:, it], rgb_out=rgb, feat_out=featc, feat_channel=it, label=gt,
                          label_plane_out=planes[:, 5], label_small_out=lsmall)
    ops.bilinear_lift(imgs[it % 2], (64, 64), out=small)
    ops.bilinear_lift(small, (H, W), out=up)
    ops.plms_step_noise(xs[0], xs[1:5], 4, 1.0, -0.1, 0.5, xs[5], 0.5, seed=1, out=o1, noisy_out=o2)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
