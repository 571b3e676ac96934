"""The lift+argmax kernels (row form, column form) and the uint16 paint + histogram at the bench shape, three
iterations, the last one between cudaProfilerStart/Stop (ncu --profile-from-start off; compute-sanitizer)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ldiffusion_b200 import _cabi, ops

torch.manual_seed(0)
dev, B, H, W, K = "cuda", 8, 1024, 1024, 11
lib = _cabi.lib()
feat = torch.randn(B, 256, 32, 32, device=dev).bfloat16()
w = (torch.randn(K, 256, device=dev) / 16).bfloat16()
logits = ops.head_logits(feat, w, None)
smooth = torch.nn.functional.interpolate(torch.randn(B, K, 4, 4, device=dev) * 3, size=(32, 32), mode="bilinear").contiguous()
mask = torch.empty(B, H, W, dtype=torch.uint8, device=dev)
gt = torch.randint(0, K, (B, H, W), dtype=torch.uint8, device=dev)
C = torch.zeros(K + 1, K, dtype=torch.int64, device=dev)
inst16 = torch.from_numpy(torch.randint(0, 801, (B, H, W)).numpy().astype("uint16")).to(dev)
lut = torch.randint(0, K, (B, 801), dtype=torch.uint8, device=dev)
for it in range(3):
    if it == 2:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
    for variant in (0, 1):
        lib.ldiff_tune(_cabi.TUNE_ARGMAX_VARIANT, variant)
        ops._lift_argmax(logits, mask)
        ops._lift_argmax(smooth, mask)
    lib.ldiff_tune(_cabi.TUNE_ARGMAX_VARIANT, 0)
    ops.lut_paint_hist(inst16, lut, gt, K, out=C, mask_out=mask)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
ops.check_status(dev)
print("done")
