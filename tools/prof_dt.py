"""ncu target: the decode tail at one pipeline shape (env TMA), gray-only and gray + step feature."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ldiffusion_b200 import _cabi, ops

lib = _cabi.lib()
lib.ldiff_tune(_cabi.TUNE_DECODE_TAIL_TMA, int(os.environ.get("TMA", "11")))
dev = "cuda"
B, H, W = 8, 1024, 1024
imgs = [torch.empty(B, 3, H, W, device=dev, dtype=torch.bfloat16).uniform_(-1.2, 1.2) for _ in range(3)]
planes = torch.empty(B, 6, H, W, dtype=torch.uint8, device=dev)
featc = torch.empty(B, 5, 64, 64, dtype=torch.bfloat16, device=dev)
for it in range(3):
    if it == 2:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
    ops.decode_tail_gray(imgs[it], want_rgb=False, gray_out=planes[:, it])
    ops.decode_tail_fused(imgs[it], planes[:, it], feat_out=featc, feat_channel=it)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
