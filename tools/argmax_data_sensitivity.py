import sys, os
sys.path.insert(0, "/root/repo")
import torch, numpy as np
from ldiffusion_b200 import ops
from ldiffusion_b200.pipeline import HotPath, HotPathInputs, synth_inputs
dev = torch.device("cuda")
B,H,W,K,N=8,1024,1024,11,5
hp = HotPath(B,H,W,K,N,device=dev,seed=1234)
for seed in (1234, 4321, 100):
    hs = synth_inputs(B,H,W,K,N,dtype=torch.bfloat16,device="cpu",seed=seed)
    feat = hs.head_feat.to(dev)
    ops._head_logits(feat, hp.head_w, hp.head_b, hp.logits)
    torch.cuda.synchronize()
    lg = hp.logits
    ref = torch.nn.functional.conv2d(feat.float(), hp.head_w.float()[:, :, None, None])
    print(seed, "nan", torch.isnan(lg).sum().item(), "maxdiff", (lg-ref).abs().max().item(), "std", lg.std().item())
    # time lift_argmax
    for _ in range(3): ops._lift_argmax(hp.logits, hp.mask_tissue)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): ops._lift_argmax(hp.logits, hp.mask_tissue)
    e1.record(); torch.cuda.synchronize()
    print("  lift_argmax us", e0.elapsed_time(e1)/20*1e3)
    # time with reference logits (no tc)
    hp.logits.copy_(ref)
    e0.record()
    for _ in range(20): ops._lift_argmax(hp.logits, hp.mask_tissue)
    e1.record(); torch.cuda.synchronize()
    print("  lift_argmax(ref logits) us", e0.elapsed_time(e1)/20*1e3)
    # histogram of classes
    print("  class hist", torch.bincount(hp.mask_tissue.view(-1).long(), minlength=K).tolist())
