"""lift_argmax alone at the bench shape (8 x 11 x 32 x 32 logits -> 8 x 1024 x 1024 mask), for one
LDIFF_ARGMAX_VARIANT per process (the knob is read once); random and smooth logits."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ldiffusion_b200 import ops
from kbench import timeit

feat = torch.randn(8, 256, 32, 32, device="cuda").bfloat16()
w = (torch.randn(11, 256, device="cuda") / 16).bfloat16()
logits = ops.head_logits(feat, w, None)
smooth = torch.nn.functional.interpolate(torch.randn(8, 11, 4, 4, device="cuda") * 3, size=(32, 32), mode="bilinear")
mask = torch.empty(8, 1024, 1024, dtype=torch.uint8, device="cuda")
res = []
for name, lg in (("random", logits), ("smooth", smooth.contiguous())):
    res.append(f"{name} {timeit(lambda i: ops._lift_argmax(lg, mask), 1):.2f} us")
print("LDIFF_ARGMAX_VARIANT=" + os.environ.get("LDIFF_ARGMAX_VARIANT", "default"), " ".join(res), flush=True)
