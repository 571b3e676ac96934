"""Oracle for stage a-3: decode tail -> uint8 image -> PIL gray -> pixel vectors.

Test infrastructure only (see ``oracle/__init__.py``).

Reference call sites: ``pixel_latent_vector.py:80-93``, ``segmentor.py:105-108``,
``:446-448``, ``:528-530``, ``utils.py:203-205``.  The helpers they call are
diffusers 0.34.0's ``decode_latents`` tail and ``numpy_to_pil`` (third-party,
not vendored; restated from the published code)::

    image = (image / 2 + 0.5).clamp(0, 1)               # in the VAE's dtype
    image = image.cpu().permute(0, 2, 3, 1).float().numpy()
    images = (images * 255).round().astype("uint8")     # numpy, half-to-even
    pil = Image.fromarray(image)                        # RGB

followed by ``PIL.Image.convert("L")`` (ITU-R 601-2 luma in 16.16 fixed point:
``(19595 R + 38470 G + 7471 B + 0x8000) >> 16``) and the per-pixel stack
``[g_0 .. g_{n-1}, label]`` (``pixel_latent_vector.py:89-93``).
"""
import numpy as np
import torch
from PIL import Image


def decode_tail_chain(image):
    """image: [B,3,H,W] fp32 or bf16 (VAE decoder output) -> uint8 [B,H,W,3]."""
    x = (image / 2 + 0.5).clamp(0, 1)
    x = x.cpu().permute(0, 2, 3, 1).float().numpy()
    return (x * 255).round().astype("uint8")


def gray_chain(rgb_u8):
    """uint8 [B,H,W,3] -> uint8 [B,H,W] through PIL, as pixel_latent_vector.py:85-86."""
    return np.stack([np.array(Image.fromarray(im).convert("L")) for im in rgb_u8])


def gray_spec(rgb_u8):
    """Integer restatement of PIL's RGB->L conversion (bit-exact target)."""
    r = rgb_u8[..., 0].astype(np.uint32)
    g = rgb_u8[..., 1].astype(np.uint32)
    b = rgb_u8[..., 2].astype(np.uint32)
    return ((19595 * r + 38470 * g + 7471 * b + 0x8000) >> 16).astype(np.uint8)


def pixel_vectors_chain(decoded_steps, label):
    """pixel_latent_vector.py:84-93 without the Python dict: decoded_steps is a
    list of n tensors [1,3,H,W]; label uint8 [H,W].  Returns uint8 [H,W,n+1]
    whose (i,j) entry is ``[g_0[i,j] .. g_{n-1}[i,j], label[i,j]]``."""
    grays = [gray_chain(decode_tail_chain(d))[0] for d in decoded_steps]
    return np.stack(grays + [np.asarray(label, dtype=np.uint8)], axis=-1)


def pixel_vectors_loop(decoded_steps, label):
    """The literal double loop of pixel_latent_vector.py:89-93 (small cases only)."""
    grays = [gray_chain(decode_tail_chain(d))[0] for d in decoded_steps]
    h, w = grays[0].shape
    lab = np.asarray(label)
    out = {}
    for i in range(h):
        for j in range(w):
            v = [grays[k][i, j] for k in range(len(grays))]
            v.append(lab[i, j])
            out[(i, j)] = v
    return out


def model_input_chain(rgb_u8, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)):
    """segmentor.py:501-510 / :533 applied to the decoded PIL image when it already is 1024x1024:
    torchvision ``ToTensor`` (uint8 HWC -> float CHW / 255) then ``Normalize(mean, std)``.
    rgb_u8: uint8 [B,H,W,3] -> float32 [B,3,H,W]."""
    from torchvision import transforms
    tf = transforms.Compose([transforms.ToTensor(), transforms.Normalize(mean=list(mean), std=list(std))])
    return torch.stack([tf(Image.fromarray(im)) for im in rgb_u8])
