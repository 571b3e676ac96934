import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ldiffusion_b200 import ops
feat = torch.randn(8, 256, 32, 32, device="cuda").bfloat16()
w = (torch.randn(11, 256, device="cuda") / 16).bfloat16()
logits = ops.head_logits(feat, w, None)
for _ in range(4):
    m = ops.lift_argmax(logits, (1024, 1024))
torch.cuda.synchronize()
