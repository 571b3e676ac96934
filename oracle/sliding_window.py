"""Oracle for the widening row N2: nnU-Net sliding-window accumulate / TTA merge / export.

Test infrastructure only (see ``oracle/__init__.py``).  Restates, with the literal eager
torch ops on fp16 CPU tensors, ``predict_from_raw_data.py:530-545`` (mirror TTA),
``:547-589`` (gaussian-weighted accumulation and normalisation in torch.half) and
``label_handling.py:128-173`` (``logits.float()`` -> softmax(0) -> argmax(0)) of the
vendored nnU-Net v2.6.2.  ``compute_gaussian`` / ``compute_steps_for_sliding_window`` are
pinned against the vendored functions by ``tests/golden/nnunet_sliding.npz``.
"""
import itertools

import torch


def tta_merge_chain(network, x, mirror_axes):
    """predict_from_raw_data.py:530-545 (network output taken as half, as under autocast)."""
    prediction = network(x).half()
    if mirror_axes is not None:
        combos = [c for i in range(len(mirror_axes)) for c in itertools.combinations([m + 2 for m in mirror_axes], i + 1)]
        for axes in combos:
            prediction += torch.flip(network(torch.flip(x, (*axes,))).half(), (*axes,))
        prediction /= (len(combos) + 1)
    return prediction


def accumulate_chain(tile_predictions, slicers, gaussian, num_heads, image_shape):
    """predict_from_raw_data.py:563-580: returns (predicted_logits half [K,H,W] after the
    normalisation, n_predictions)."""
    predicted_logits = torch.zeros((num_heads, *image_shape), dtype=torch.half)
    n_predictions = torch.zeros(image_shape, dtype=torch.half)
    for prediction, sl in zip(tile_predictions, slicers):
        predicted_logits[sl] += (prediction * gaussian if gaussian is not None else prediction)
        n_predictions[sl[1:]] += (gaussian if gaussian is not None else 1)
    predicted_logits /= n_predictions
    return predicted_logits, n_predictions


def export_chain(predicted_logits):
    """label_handling.py:128-173 (no regions): logits.float() -> softmax(0) -> argmax(0)."""
    return torch.softmax(predicted_logits.float(), 0).argmax(0)
