"""Oracle for stage a-5: per-pixel classifier head + argmax.

Test infrastructure only (see ``oracle/__init__.py``).

Tissue form: ``TissueSegNet.decoder[-1]`` 1x1 conv 256->K (``conductor.py:127``)
-> ``F.interpolate(out, size=x.shape[2:], mode='bilinear',
align_corners=False)`` (``conductor.py:135``) -> ``argmax(softmax(out, 1), 1)``
(``segmentor.py:536``; same idiom at ``predict_from_raw_data.py:918``).

Cell form: ``Linear(256, K)`` on per-instance features (``conductor.py:218``)
-> ``softmax[:, 1:]`` -> ``topk(1) + 1`` (``:219-221``) -> paint a one-hot mask
per instance (``:224-231``) -> the same ``argmax(softmax)`` at
``segmentor.py:536``.

Exactness.  The contraction (1x1 conv / Linear) is floating point with an
implementation-defined summation order (cuDNN / MKL in the reference, tensor
cores in the build): it is compared within 1e-3 relative.  Everything after the
low-resolution logits is pinned: the lift is ``oracle.bilinear.lift_spec``
(bit-identical to ATen-CPU for up-sampling) and the decision rule is
``softmax_argmax_spec`` below.  ``argmax(softmax(x))`` differs from
``argmax(x)`` only where rounding merges two probabilities (then the lower
index wins); ATen's vectorised ``exp`` is not correctly rounded, so the literal
chain's behaviour at such near-ties is implementation noise.  The spec fixes
it: ``e_k = fl32(exp_f64(fl32(x_k - max)))``, ``S`` = sequential fp32 sum over
k, ``p_k = fl32(e_k / S)``, first index of the maximum.  Tests report how many
pixels of the chain disagree with the spec and check that every one of them is
a near-tie (top-2 logit gap <= 4 ulp); on the seeded configs the count is 0.
"""
import numpy as np
import torch
import torch.nn.functional as F

from ._fp import div32, sub32
from .bilinear import lift_chain, lift_spec


# ---------------------------- chain tier ----------------------------------

def conv1x1_chain(feat, weight, bias):
    """conductor.py:127: nn.Conv2d(256, K, 1).  feat [B,C,h,w], weight [K,C]."""
    return F.conv2d(feat, weight[:, :, None, None], bias)


def head_argmax_chain(feat, weight, bias, size):
    """conductor.py:127,135 + segmentor.py:536.  Returns (mask int64 [B,H,W],
    low-res logits)."""
    logits_lr = conv1x1_chain(feat, weight, bias)
    out = lift_chain(logits_lr, size)
    return torch.argmax(torch.softmax(out, dim=1), dim=1), logits_lr


def lift_argmax_chain(logits_lr, size):
    out = lift_chain(logits_lr, size)
    return torch.argmax(torch.softmax(out, dim=1), dim=1)


def cell_classify_chain(inst_feats, weight, bias):
    """conductor.py:218-221.  inst_feats [N,256] -> class ids int64 [N] in 1..K-1."""
    logits = F.linear(inst_feats, weight, bias)
    probs = F.softmax(logits, dim=1)[:, 1:]
    _, lab = torch.topk(probs, k=1, dim=1)
    return (lab + 1).reshape(-1), logits


def cell_paint_chain(inst_map, inst_ids, class_ids, num_classes):
    """conductor.py:224-231 + segmentor.py:536: accumulate one-hot planes per
    instance, then argmax(softmax).  inst_map int [H,W]; returns int64 [H,W].
    (Allocates [1,K,H,W] per instance like the reference: small cases only.)"""
    H, W = inst_map.shape
    final = torch.zeros((1, num_classes, H, W))
    for inst, cls in zip(inst_ids, class_ids):
        m = torch.as_tensor(np.asarray(inst_map) == int(inst), dtype=torch.float32)
        onehot = F.one_hot(torch.tensor(int(cls)), num_classes=num_classes).float()[:, None, None]
        final = final + onehot * m.unsqueeze(0)
    return torch.argmax(torch.softmax(final, dim=1), dim=1)[0]


# ---------------------------- spec tier ------------------------------------

def softmax_argmax_spec(x):
    """x: fp32 [B,K,...] -> uint8 [B,...] by the pinned rule in the module doc."""
    x = np.asarray(x, np.float32)
    m = x.max(axis=1, keepdims=True)
    e = np.exp(sub32(x, m).astype(np.float64)).astype(np.float32)
    s = np.zeros_like(e[:, 0])
    for k in range(x.shape[1]):
        s = (s + e[:, k]).astype(np.float32)
    p = div32(e, s[:, None])
    return np.argmax(p, axis=1).astype(np.uint8)        # np.argmax: first maximum


def lift_argmax_spec(logits_lr, size):
    """fp32 low-res logits [B,K,h,w] -> uint8 mask [B,H,W]."""
    return softmax_argmax_spec(lift_spec(logits_lr, size))


def top2_gap_ulps(x):
    """Per pixel: (x_max - x_second) / ulp(x_max), for the near-tie report."""
    x = np.asarray(x, np.float32)
    srt = np.sort(x, axis=1)
    gap = (srt[:, -1] - srt[:, -2]).astype(np.float64)
    return gap / np.spacing(np.abs(srt[:, -1]).astype(np.float32)).astype(np.float64)


def cell_lut_spec(inst_ids, class_ids, n_entries):
    """LUT form of the painting loop: lut[inst] = class, lut[0] = 0 (background
    and instances the reference skips stay 0)."""
    lut = np.zeros(n_entries, np.uint8)
    lut[np.asarray(inst_ids, np.int64)] = np.asarray(class_ids, np.uint8)
    lut[0] = 0
    return lut


def cell_paint_spec(inst_map, lut):
    return lut[np.asarray(inst_map, np.int64)]
