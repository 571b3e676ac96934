"""Launch each hot-path kernel a few times at the BASELINE config-2 shapes (for ncu)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ldiffusion_b200 import ops
from ldiffusion_b200.pipeline import synth_inputs

dev = "cuda"
dt = torch.bfloat16
inp = synth_inputs(8, 1024, 1024, 11, 2, dtype=dt, device=dev, n_instances=800)
planes = torch.empty(8, 6, 1024, 1024, dtype=torch.uint8, device=dev)
rgb = torch.empty(8, 1024, 1024, 3, dtype=torch.uint8, device=dev)
w = (torch.randn(11, 256, device=dev) / 16).to(dt)
src = torch.randn(8, 3, 64, 64, device=dev).to(dt)
up = torch.empty(8, 3, 1024, 1024, device=dev, dtype=dt)
fc = torch.empty(8, 5, 64, 64, device=dev, dtype=dt)
lut = torch.randint(0, 11, (8, 801), device=dev, dtype=torch.uint8)
C = torch.zeros(12, 11, dtype=torch.int64, device=dev)
x = inp.latents
o = torch.empty_like(x)
sc = torch.rand(8, 1, 128, 128, device=dev).to(dt)
cw = (torch.randn(11, 256, device=dev) / 16).to(dt)
ids = torch.arange(1, 801, dtype=torch.int32, device=dev)
for it in range(2):
    ops.decode_tail_gray(inp.decoded[it % 2], want_rgb=False, gray_out=planes[:, it])
    ops.decode_tail_gray(inp.decoded[it % 2], rgb_out=rgb, gray_out=planes[:, it])
    ops.bilinear_lift(src, (1024, 1024), out=up)
    ops.bilinear_lift(inp.decoded[it % 2], (64, 64), out=fc, out_channel=it, gray=True)
    logits = ops.head_logits(inp.head_feat, w, None)
    mask = ops.lift_argmax(logits, (1024, 1024))
    ops.lut_paint(inp.inst_map, lut)
    ops.confusion_hist(mask.view(-1), inp.gt.view(-1), 11, out=C)
    ops.laplace_qsample(x, 0.9, seed=1, offset=it, out=o)
    ops.plms_step(x, [inp.eps[0], inp.eps[1], inp.eps[0], inp.eps[1]], 4, 1.01, 0.02, 0.9, out=o)
    ops.laplace_qsample_map(x, sc, seed=1, offset=it, x_mul=0.18215, out=o)
    ops.scaled_residual(x, inp.eps[0], sc, out_div=0.18215, out=o)
    ops.cell_classify(inp.inst_feats, cw, None, ids, 801)
torch.cuda.synchronize()
print("done")
