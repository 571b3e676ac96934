"""Widening N3: the pixel-contrastive InfoNCE term of the warm-up loss.

Drop-in for ``InfoNceLoss.compute_contrastive_loss(features, labels)`` (``model/loss.py:44-109``),
the caller of the training-path feature tensor (``ldiffusion.py:252``).  The reference samples
(anchor, positive, 1024 negatives) triples with ``torch.randperm`` / ``torch.randint`` on the
host and then runs a Python loop of 1 x C by C x 1025 matmuls + ``cross_entropy`` per triple; here
ALL triples are evaluated by one warp each in a single launch, with a hand-written backward, and
the sampling itself runs in one more launch (``sample_contrastive_pairs_device``: same procedure,
counter-based randomness, no ``.item()`` round trips); ``sample_contrastive_pairs`` keeps the
reference's host procedure with a ``torch.Generator`` for comparison.

The VGG content term (``loss.py:19-42``) is a torchvision network: out of scope.
"""
from typing import Optional, Tuple

import torch

from . import ops

Pairs = Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]


def sample_contrastive_pairs(labels: torch.Tensor, num_negatives: int = 1024,
                             generator: Optional[torch.Generator] = None) -> Optional[Pairs]:
    """The sampling of loss.py:64-87.  labels: integer [B,1,h,w] (any device; sampling runs on the
    host like the reference's ``.item()`` / ``.tolist()`` loop).  Returns int32 CPU tensors
    (pair_batch [A], anchor [A], positive [A], negatives [A, num_negatives]) or None."""
    B = labels.shape[0]
    lab = labels.reshape(B, -1).cpu()
    pb, pa, pq, pn = [], [], [], []
    for b in range(B):
        label = lab[b]
        for lbl in torch.unique(label):
            mask = label == lbl
            pos_idx = torch.nonzero(mask).squeeze(-1)
            neg_idx = torch.nonzero(~mask).squeeze(-1)
            if len(pos_idx) > 1 and len(neg_idx) > num_negatives:
                sampled = torch.randperm(len(pos_idx), generator=generator)[:max(1, int(0.01 * len(pos_idx)))]
                for idx in sampled:
                    anchor = pos_idx[idx].item()
                    pool = pos_idx[pos_idx != anchor]
                    if len(pool) == 0:
                        continue
                    pos = pool[torch.randint(0, len(pool), (1,), generator=generator)].item()
                    neg = neg_idx[torch.randperm(len(neg_idx), generator=generator)[:num_negatives]]
                    pb.append(b); pa.append(anchor); pq.append(pos); pn.append(neg)
    if not pa:
        return None
    i32 = lambda v: torch.tensor(v, dtype=torch.int32)                     # noqa: E731
    return i32(pb), i32(pa), i32(pq), torch.stack(pn).to(torch.int32)


def sample_contrastive_pairs_device(labels: torch.Tensor, num_negatives: int = 1024, seed: int = 0,
                                    offset: int = 0):
    """The sampling of loss.py:64-87 in one launch (``ldiff_infonce_sample``).  labels: integer
    [B,1,h,w] on the GPU, values < 32, h*w <= 16384.  Returns int32 device tensors (pair_batch
    [B*cap], anchor, positive, negatives [B*cap, N], n_valid [B]); slots a class did not fill carry
    pair_batch = -1.  Deterministic in (seed, offset); nothing is read back to the host."""
    if not labels.is_cuda:
        raise ops.LdiffError("ldiff operators run on CUDA tensors only (no CPU fallback)")
    B = labels.shape[0]
    lab = labels.reshape(B, -1)
    lab = (lab if lab.dtype == torch.uint8 else lab.clamp(0, 255).to(torch.uint8)).contiguous()
    hw = lab.shape[1]
    cap = hw // 100 + 32
    dev = lab.device
    i32 = lambda *s: torch.empty(s, dtype=torch.int32, device=dev)         # noqa: E731
    pb, pa, pq, neg, nv = i32(B * cap), i32(B * cap), i32(B * cap), i32(B * cap, int(num_negatives)), i32(B)
    ops._infonce_sample(lab, int(num_negatives), cap, int(seed), int(offset), pb, pa, pq, neg, nv,
                        ops.status_word(dev))
    return pb, pa, pq, neg, nv


class _InfoNce(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, pb, pa, pq, neg, count, temperature):
        A = pa.numel()
        loss = torch.empty(A, dtype=torch.float32, device=feat.device)
        lse = torch.empty(A, dtype=torch.float32, device=feat.device)
        ops._infonce_forward(feat, pb, pa, pq, neg, loss, lse, float(temperature))
        ctx.save_for_backward(feat, pb, pa, pq, neg, lse, count)
        ctx.temperature = float(temperature)
        return loss.sum() / count                              # == mean over the triples (loss.py:109)

    @staticmethod
    def backward(ctx, grad_out):
        feat, pb, pa, pq, neg, lse, count = ctx.saved_tensors
        grad = torch.zeros_like(feat)
        gscale = (grad_out.to(torch.float32) / count).reshape(1).contiguous()
        ops._infonce_backward(feat, pb, pa, pq, neg, lse, gscale, grad, ctx.temperature)
        return grad, None, None, None, None, None, None


def pixel_contrastive_loss(features: torch.Tensor, labels: torch.Tensor, temperature: float = 0.5,
                           num_negatives: int = 1024, pairs: Optional[Pairs] = None,
                           generator: Optional[torch.Generator] = None, sampler: str = "device",
                           seed: int = 0, offset: int = 0) -> torch.Tensor:
    """``InfoNceLoss.compute_contrastive_loss`` (loss.py:44-109).  features: fp32 [B,n,h,w] on the
    GPU (n <= 16 per-step gray channels); labels [B,1,h,w].  ``pairs`` injects the index sets
    (parity runs).  Otherwise they are drawn with the reference's rules, by ``sampler="device"``
    (one launch, ``seed``/``offset`` — advance ``offset`` every training step) or
    ``sampler="host"`` (the reference's own torch.randperm / randint procedure, ``generator``)."""
    if not features.is_cuda:
        raise ops.LdiffError("ldiff operators run on CUDA tensors only (no CPU fallback)")
    if features.dim() != 4 or features.shape[1] > 16:
        raise ValueError("features must be [B,n,h,w] with n <= 16")
    if sampler not in ("device", "host"):
        raise ValueError("sampler must be 'device' or 'host'")
    dev = features.device
    feat = features.to(torch.float32).contiguous()
    if pairs is None and sampler == "device":
        pb, pa, pq, neg, nv = sample_contrastive_pairs_device(labels.to(dev), num_negatives, seed, offset)
        # no valid triple: every slot is skipped, the sum is 0 and so is the gradient (loss.py:106-107)
        count = nv.sum().clamp(min=1).to(torch.float32)
        return _InfoNce.apply(feat, pb, pa, pq, neg, count, temperature)
    if pairs is None:
        pairs = sample_contrastive_pairs(labels, num_negatives, generator)
    if pairs is None:                                                      # loss.py:106-107
        return torch.tensor(0.0, requires_grad=True, device=dev)
    pb, pa, pq, neg = (t.to(dev, torch.int32).contiguous() for t in pairs)
    count = torch.tensor(float(pa.numel()), device=dev)
    return _InfoNce.apply(feat, pb, pa, pq, neg, count, temperature)
