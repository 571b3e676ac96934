"""Oracle for the widening row N3: pixel-contrastive InfoNCE (test infrastructure only).

Restates ``InfoNceLoss.compute_contrastive_loss`` (``model/loss.py:44-109``) with the sampled
index sets injected (the reference draws them with unseeded ``torch.randperm``): the literal
per-triple loop of ``:92-104`` — 1 x C by C x (1+N) matmuls, ``/ temperature``, ``torch.cat``,
``F.cross_entropy(logits, target=0)`` — and the final mean (``:109``).  ``model/loss.py`` itself
cannot be imported here (it imports diffusers at the top), so this tier is pinned only against
torch's own ``matmul`` / ``cross_entropy``.
"""
import torch
import torch.nn.functional as F


def contrastive_loss_chain(features, pairs, temperature=0.5):
    """features: float [B,n,h,w] (may require grad); pairs: (pair_batch, anchor, pos, neg[A,N])."""
    B, n, h, w = features.shape
    feats = features.view(B, n, -1).permute(0, 2, 1)                       # [B, h*w, n]  (loss.py:59)
    pb, pa, pq, neg = pairs
    total, count = 0.0, 0
    for i in range(len(pa)):
        feat = feats[int(pb[i])]
        anchor = feat[int(pa[i])].unsqueeze(0)
        positive = feat[int(pq[i])].unsqueeze(0)
        negatives = feat[neg[i].long()]
        pos_sim = torch.matmul(anchor, positive.t()) / temperature
        neg_sim = torch.matmul(anchor, negatives.t()) / temperature
        logits = torch.cat([pos_sim, neg_sim], dim=-1)
        total = total + F.cross_entropy(logits, torch.tensor([0], dtype=torch.long))
        count += 1
    if count == 0:
        return torch.tensor(0.0, requires_grad=True)
    return total / count
