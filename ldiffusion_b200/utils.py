"""Dataset-preparation callers of the sampling loop (``utils.py:155-208`` of the reference).

``copy_or_convert_image`` is the fourth copy of the one-step sampling loop: image ->
VAE encode -> ``set_timesteps(1)`` -> UNet -> ``scheduler.step`` -> decode -> PIL.  Here it runs
through ``Segmentor._sample_and_decode`` (one fused scheduler launch, decode tail on the device).
``convert_and_save_label`` is the label table remap (see ``dataset.py``).
"""
import shutil

import numpy as np
import torch
from PIL import Image

from .dataset import label_lut_numpy


def check_images_same_size(image_paths):
    """utils.py:155-163."""
    sizes = set()
    for path in image_paths:
        with Image.open(path) as img:
            sizes.add(img.size)
            if len(sizes) > 1:
                return False
    return True


def convert_and_save_label(lbl, dst_path, mapping):
    """utils.py:165-173: gray-level label image -> class indices by ``mapping`` -> PNG."""
    arr = np.array(lbl)
    if arr.dtype != np.uint8:
        converted = np.zeros_like(arr, dtype=np.uint8)            # the reference's loop, for exotic modes
        for k, v in mapping.items():
            converted[arr == k] = v
    else:
        converted = label_lut_numpy(mapping)[arr]
    Image.fromarray(converted).save(dst_path)


@torch.no_grad()
def copy_or_convert_image(img, src_path, dst_path, pipeline=None, unet=None, use_diffusion=True):
    """utils.py:176-208: one-step L-Diffusion reconstruction of ``img`` saved as PNG, or a plain copy."""
    if not use_diffusion:
        shutil.copy(src_path, dst_path)
        return
    from .scheduler import LaplacePLMSScheduler
    from .segmentor import Segmentor, _to_tensor_1024
    seg = Segmentor(None, None, "tissue", 1)
    image = _to_tensor_1024(img, seg.device, normalize=True)                              # :180-186
    if not isinstance(pipeline.scheduler, LaplacePLMSScheduler):
        pipeline.scheduler = LaplacePLMSScheduler()
    # :192-197 (a fresh random text projection per call, as in the reference; sized to what the UNet
    # accepts — the reference's hard-coded 1280 only works through its wrapper's fallback)
    linear_layer = torch.nn.Linear(pipeline.text_encoder.config.hidden_size,
                                   unet.config.cross_attention_dim).to(seg.device)
    ids = torch.tensor(pipeline.tokenizer(["A pathological slide"] * 1)["input_ids"], device=seg.device)
    text = linear_layer(pipeline.text_encoder(ids)["last_hidden_state"].float()).detach()
    rgb = seg._sample_and_decode(image, pipeline, unet, pipeline.vae, text, num_steps=1)   # :188-206
    Image.fromarray(rgb[0].cpu().numpy()).save(dst_path)                                  # :206-207
