"""Segmentor heads: per-pixel classifier + argmax (tissue) and instance
classifier + LUT painting (cell).

Drop-in for the tails of ``TissueSegNet.forward`` (``conductor.py:127,135``) and
``CellSegClassifier.forward`` (``conductor.py:218-231``) followed by the
``argmax(softmax(out, 1), 1)`` of ``segmentor.py:536``.  The backbones (ConvNeXt,
ResNet-152, Cellpose) are library calls outside this package.
"""
from typing import Optional

import torch

from . import ops


class TissueHead(torch.nn.Module):
    """``nn.Conv2d(256, K, 1)`` + bilinear lift to the input size + argmax, fused:
    writes only the uint8 mask.  ``forward`` keeps the reference's ``{"out": ...}``
    contract when ``return_logits`` is set (low-resolution logits, not the lift)."""

    def __init__(self, num_classes: int, in_channels: int = 256, dtype=torch.bfloat16, device="cuda"):
        super().__init__()
        conv = torch.nn.Conv2d(in_channels, num_classes, 1)
        self.weight = torch.nn.Parameter(conv.weight.detach().reshape(num_classes, in_channels).to(device, dtype))
        self.bias = torch.nn.Parameter(conv.bias.detach().to(device, torch.float32))
        self.num_classes = num_classes

    @torch.no_grad()
    def forward(self, feat: torch.Tensor, size, return_logits: bool = False):
        return ops.head_argmax(feat, self.weight, self.bias, size, return_logits=return_logits)


def tissue_mask(feat, weight, bias, size):
    """feat [B,256,h,w] -> uint8 mask [B,H,W]."""
    return ops.head_argmax(feat, weight, bias, size)


def cell_mask(inst_map: torch.Tensor, inst_feats: torch.Tensor, weight: torch.Tensor,
              bias: Optional[torch.Tensor], inst_ids: torch.Tensor, lut_size: Optional[int] = None):
    """conductor.py:218-231 + segmentor.py:536 for one image or a batch.

    inst_map int32 or uint16 [H,W] (Cellpose instance ids, 0 = background); inst_feats
    [N,256] pooled features of the instances the reference keeps; inst_ids [N].
    Instances the reference skips simply have no LUT entry (class 0).
    """
    if lut_size is None:
        lut_size = int(inst_ids.max().item()) + 1 if inst_ids.numel() else 1
    lut = ops.cell_classify(inst_feats, weight, bias, inst_ids, lut_size)
    return ops.lut_paint(inst_map, lut)
