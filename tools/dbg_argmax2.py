import sys
sys.path.insert(0, "/root/repo")
import torch, torch.nn.functional as F
from ldiffusion_b200.pipeline import HotPath, synth_inputs
dev = torch.device("cuda")
B,H,W,K,N=8,1024,1024,11,5
hp = HotPath(B,H,W,K,N,device=dev,seed=1234)
for seed in (1234, 4321):
    hs = synth_inputs(B,H,W,K,N,dtype=torch.bfloat16,device="cpu",seed=seed)
    feat = hs.head_feat.to(dev)
    ref = F.conv2d(feat.float(), hp.head_w.float()[:, :, None, None])
    for b in range(B):
        x = F.interpolate(ref[b:b+1], size=(H,W), mode="bilinear", align_corners=False)[0]
        top = torch.topk(x, 2, dim=0).values
        gap = top[0]-top[1]
        amb = gap <= 1e-5
        # events per thread region: columns pairs x//2, bands y//32
        ys, xs = torch.nonzero(amb, as_tuple=True)
        key = (ys//32)*10000 + xs//2
        uniq, cnt = torch.unique(key, return_counts=True)
        print(seed, b, "amb", int(amb.sum()), "threads", len(uniq), "max per thread", int(cnt.max()) if len(cnt) else 0, "absmax", float(x.abs().max()))
