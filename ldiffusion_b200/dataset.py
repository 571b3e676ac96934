"""Label tables of the reference (``dataset.py:10-63``) as 256-entry look-up tables.

The reference remaps a gray-level label image with one boolean mask per table entry
(``convert_labels`` ``dataset.py:48-63``, ``map_mask`` ``:42-46``, ``utils.convert_and_save_label``
``utils.py:165-173``).  Every one of them is the same operation: ``out = lut[gray]`` with
``lut[k] = v`` for the table's entries and 0 elsewhere.  ``label_lut`` builds that table once;
on the device it is the ``gt_lut`` argument of ``ldiff_confusion_hist`` (the remap is fused into
the histogram's gt read), on the host it is one numpy gather.
"""
import numpy as np
import torch
from PIL import Image

# dataset.py:10-40, verbatim values
pixel_to_label = {0: 0, 100: 1, 150: 2, 50: 3, 200: 4, 250: 5, 255: 6}
pixel_to_label_cell = {0: 0, 25: 1, 50: 2, 75: 3, 100: 4, 125: 5, 150: 6, 175: 7, 200: 8, 225: 9, 250: 10}
ID_TO_CLASS = {0: 0, 60: 1, 120: 2, 180: 3, 255: 0}


def label_lut_numpy(mapping) -> np.ndarray:
    """uint8 [256]: lut[k] = v for (k, v) in mapping, 0 elsewhere — what the reference's
    ``zeros_like`` + ``out[arr == k] = v`` loop computes for a uint8 image."""
    lut = np.zeros(256, dtype=np.uint8)
    for k, v in mapping.items():
        if not (0 <= int(k) <= 255 and 0 <= int(v) <= 255):
            raise ValueError(f"label table entry {k}: {v} does not fit a uint8 gray level")
        lut[int(k)] = int(v)
    return lut


def label_lut(mapping, device="cuda") -> torch.Tensor:
    """The table as the device tensor ``confusion_hist(..., gt_lut=)`` / ``confusion_matrix`` take."""
    return torch.from_numpy(label_lut_numpy(mapping)).to(device)


def level_table(level: str):
    if level == "tissue":
        return pixel_to_label
    if level == "cell":
        return pixel_to_label_cell
    raise ValueError("Unsupported level. Use 'tissue' or 'cell'.")


def convert_labels(img_path, level):
    """dataset.py:48-63 -> uint8 [H,W] class map."""
    table = level_table(level)
    img_array = np.array(Image.open(img_path).convert("L"), dtype=np.uint8)
    return label_lut_numpy(table)[img_array]


def map_mask(mask_np):
    """dataset.py:42-46 -> int64 class map (values outside the table -> 0)."""
    mask_np = np.asarray(mask_np)
    inside = (mask_np >= 0) & (mask_np <= 255)
    lut = label_lut_numpy(ID_TO_CLASS).astype(np.int64)
    return np.where(inside, lut[np.clip(mask_np, 0, 255).astype(np.int64)], 0).astype(np.int64)
