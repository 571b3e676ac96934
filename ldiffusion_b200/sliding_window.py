"""Widening N2: nnU-Net's 2-D sliding-window inference tail on the tissue path.

Drop-in pieces of the vendored nnU-Net v2.6.2 (``model/nnunetv2``):

* ``compute_gaussian`` / ``compute_steps_for_sliding_window``
  (``inference/sliding_window_prediction.py:10-56``) — host-side setup, same values;
* ``SlidingWindowAccumulator`` — the gaussian-weighted ``predicted_logits[sl] += ...`` /
  ``n_predictions[sl] += ...`` / ``predicted_logits /= n_predictions`` loop of
  ``predict_from_raw_data.py:547-589`` (fp16 results arrays, as there);
* ``tta_merge`` — the mirror-TTA accumulation of ``:530-545`` in one launch;
* ``SlidingWindowAccumulator.finalize`` — normalisation fused with the export
  ``logits.float() -> softmax(0) -> argmax(0)`` (``label_handling.py:128-173``), writing the
  uint8 segmentation directly.

The network itself, the pre-processing and the file export are out of scope.
"""
import itertools
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops


# The two set-up helpers below must return exactly the values of nnU-Net v2's functions of the same name
# (nnunetv2/inference/sliding_window_prediction.py, (c) Division of Medical Image Computing, DKFZ, Apache-2.0): the
# gaussian importance map and the tile origins decide which pixels a tile contributes to and with what weight, so
# any other arithmetic would change masks.  They are restated here (golden vectors from the vendored functions pin
# them in tests/golden/nnunet_sliding.npz); the accumulation itself is this package's CUDA kernels.
def compute_gaussian(tile_size: Sequence[int], sigma_scale: float = 1. / 8, value_scaling_factor: float = 1,
                     dtype=torch.float16, device="cuda") -> torch.Tensor:
    """Importance map of one tile: a unit impulse at the tile centre blurred with sigma = sigma_scale * extent per
    axis, scaled so that its maximum is ``value_scaling_factor``, zeros (fp16 underflow at the corners) lifted to the
    smallest non-zero weight so the later division by the summed weights never sees 0."""
    from scipy.ndimage import gaussian_filter
    shape = tuple(int(t) for t in tile_size)
    impulse = np.zeros(shape)
    impulse[tuple(t // 2 for t in shape)] = 1
    blurred = gaussian_filter(impulse, [t * sigma_scale for t in shape], 0, mode="constant", cval=0)
    weights = torch.from_numpy(blurred)
    weights = (weights / torch.max(weights) * value_scaling_factor).type(dtype).to(device)
    positive = weights != 0
    weights[~positive] = torch.min(weights[positive])
    return weights


def compute_steps_for_sliding_window(image_size: Sequence[int], tile_size: Sequence[int],
                                     tile_step_size: float) -> List[List[int]]:
    """Tile origins per axis: the fewest tiles whose stride does not exceed ``tile_step_size`` tiles, spread evenly so
    that the first tile starts at 0 and the last one ends at the image border."""
    if any(i < t for i, t in zip(image_size, tile_size)):
        raise AssertionError("every image extent must be at least the tile extent")
    if not 0 < tile_step_size <= 1:
        raise AssertionError("tile_step_size must lie in (0, 1]")
    origins = []
    for extent, tile in zip(image_size, tile_size):
        count = int(np.ceil((extent - tile) / (tile * tile_step_size))) + 1
        last = extent - tile
        stride = last / (count - 1) if count > 1 else 0.0      # (one tile: its origin is 0 whatever the stride)
        origins.append([int(np.round(stride * i)) for i in range(count)])
    return origins


def tta_merge(preds: Sequence[torch.Tensor], flips: Sequence[int], out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """predict_from_raw_data.py:541-544.  preds[0]: network(x); preds[j]: network(flip(x, flips[j]))
    NOT yet flipped back (bit 0 = rows, bit 1 = columns).  Returns fp16 [K,th,tw]."""
    p0 = preds[0]
    for t in preds:
        if t.dtype != torch.float16 or t.shape != p0.shape or not t.is_contiguous():
            raise ValueError("predictions must be contiguous fp16 tensors of one shape")
    if flips[0] != 0:
        raise ValueError("the first prediction is the unflipped one")
    out = torch.empty_like(p0) if out is None else out
    ops._sw_tta_merge(list(preds), list(flips), out)
    return out


class SlidingWindowAccumulator:
    def __init__(self, num_heads: int, image_shape: Tuple[int, int], tile_size: Tuple[int, int],
                 use_gaussian: bool = True, device="cuda"):
        self.K, (self.H, self.W), self.tile = int(num_heads), tuple(image_shape), tuple(tile_size)
        self.acc = torch.zeros((self.K, self.H, self.W), dtype=torch.half, device=device)
        self.npred = torch.zeros((self.H, self.W), dtype=torch.half, device=device)
        self.gaussian = compute_gaussian(self.tile, sigma_scale=1. / 8, value_scaling_factor=10,
                                         device=device) if use_gaussian else None

    def add(self, prediction: torch.Tensor, y0: int, x0: int):
        """prediction: fp16 [K,th,tw] network output for the tile at (y0, x0)."""
        if prediction.dtype != torch.float16 or tuple(prediction.shape) != (self.K, *self.tile):
            raise ValueError("prediction must be fp16 [K, tile_h, tile_w]")
        ops._sw_accumulate(prediction.contiguous(), self.gaussian, self.acc, self.npred, int(y0), int(x0))

    def finalize(self, return_logits: bool = False):
        """-> uint8 segmentation [H,W] (and the normalised fp16 logits [K,H,W])."""
        seg = torch.empty((self.H, self.W), dtype=torch.uint8, device=self.acc.device)
        logits = torch.empty_like(self.acc) if return_logits else None
        ops._sw_finalize_argmax(self.acc, self.npred, seg, logits, ops.status_word(self.acc.device))
        return (seg, logits) if return_logits else seg


@torch.no_grad()
def predict_sliding_window(network, data: torch.Tensor, tile_size: Tuple[int, int], num_heads: int,
                           tile_step_size: float = 0.5, use_gaussian: bool = True,
                           mirror_axes: Optional[Tuple[int, ...]] = (0, 1), return_logits: bool = False):
    """2-D form of ``nnUNetPredictor.predict_sliding_window_return_logits`` + export
    (predict_from_raw_data.py:496-589, label_handling.py:128-173).  data: [C,H,W] (already >= tile)."""
    _, H, W = data.shape
    steps = compute_steps_for_sliding_window((H, W), tile_size, tile_step_size)
    sw = SlidingWindowAccumulator(num_heads, (H, W), tile_size, use_gaussian, data.device)
    combos = []
    if mirror_axes:
        combos = [c for i in range(len(mirror_axes)) for c in itertools.combinations([m + 2 for m in mirror_axes], i + 1)]
    for sy in steps[0]:
        for sx in steps[1]:
            x = data[None, :, sy:sy + tile_size[0], sx:sx + tile_size[1]]
            preds, flips = [network(x)[0].half().contiguous()], [0]
            for axes in combos:
                preds.append(network(torch.flip(x, axes))[0].half().contiguous())
                flips.append(sum((1 if a == 2 else 2) for a in axes))
            pred = tta_merge(preds, flips) if combos else preds[0]
            sw.add(pred, sy, sx)
    out = sw.finalize(return_logits=return_logits)
    ops.check_status(data.device)
    return out
