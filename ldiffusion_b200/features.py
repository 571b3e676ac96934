"""Per-pixel feature extraction: decode tail -> gray -> pixel vectors, and the
training-path bilinear lift + gray + concat.

Inference path (``pixel_latent_vector.py:58-102``): every sampling step's decoded
image becomes one uint8 gray plane; pixel (i, j)'s vector is
``[g_0[i,j] .. g_{n-1}[i,j], label[i,j]]``.  The reference builds it with a
device->host copy per step, PIL, and a Python loop over every pixel; here each
step is one kernel that writes its gray plane straight into slot i of a
preallocated ``[B, n+1, H, W]`` uint8 tensor (planar in HBM so every write is a
dense 16-byte store; ``.vectors`` exposes the reference's ``[.., H, W, n+1]``
indexing as a zero-copy permuted view).

Training path (``ldiffusion.py:224-226,240-251``): decode -> bilinear 64x64 ->
0.2989/0.5870/0.1140 gray -> concat over steps; label bilinear down + uint8
truncation; last RGB bilinear up to 1024x1024.
"""
import csv
import os
from datetime import datetime
from typing import Optional, Sequence

import torch

from . import ops


class PixelVectorBuilder:
    """Accumulates the per-step gray planes of a batch of patches."""

    def __init__(self, batch: int, height: int, width: int, num_steps: int, device):
        self.n = int(num_steps)
        self.planes = torch.empty((batch, self.n + 1, height, width), dtype=torch.uint8, device=device)
        self.rgb = None           # uint8 [B,H,W,3] of the last decoded step

    def add_step(self, i: int, decoded: torch.Tensor, want_rgb: bool = False):
        """decoded: [B,3,H,W] VAE decoder output of sampling step i (fp32 or bf16)."""
        if not 0 <= i < self.n:
            raise IndexError("step index out of range")
        rgb, _ = ops.decode_tail_gray(decoded, want_rgb=want_rgb, gray_out=self.planes[:, i])
        if want_rgb:
            self.rgb = rgb
        return rgb

    def set_label(self, label: torch.Tensor):
        """label: integer class map [B,H,W] (or [B,1,H,W]) -> last slot."""
        lab = label.reshape(self.planes.shape[0], *self.planes.shape[2:]).to(torch.uint8).contiguous()
        ops.copy_planes_u8(lab, self.planes[:, self.n])

    @property
    def vectors(self) -> torch.Tensor:
        """uint8 [B,H,W,n+1] view: vectors[b,i,j] == [g_0 .. g_{n-1}, label] of pixel (i,j)."""
        return self.planes.permute(0, 2, 3, 1)


def pixel_vectors(decoded_steps: Sequence[torch.Tensor], label: Optional[torch.Tensor] = None,
                  return_rgb: bool = False):
    """Tensor-returning core of ``pixel_latent_vector``: list of n decoded
    [B,3,H,W] tensors (+ label) -> PixelVectorBuilder."""
    d0 = decoded_steps[0]
    B, _, H, W = d0.shape
    pv = PixelVectorBuilder(B, H, W, len(decoded_steps), d0.device)
    for i, d in enumerate(decoded_steps):
        pv.add_step(i, d, want_rgb=return_rgb and i == len(decoded_steps) - 1)
    if label is not None:
        pv.set_label(label)
    else:
        pv.planes[:, pv.n].zero_()
    return pv


def feature_concat(decoded_steps: Sequence[torch.Tensor], size=(64, 64), out_dtype=None) -> torch.Tensor:
    """ldiffusion.py:240-247: per step bilinear -> gray -> concat on dim 1, written
    in place into a preallocated [B,n,h,w] tensor (no growing torch.cat)."""
    d0 = decoded_steps[0]
    out = torch.empty((d0.shape[0], len(decoded_steps), size[0], size[1]),
                      dtype=out_dtype or d0.dtype, device=d0.device)
    same = all(d.shape == d0.shape and d.dtype == d0.dtype and d.stride() == d0.stride() for d in decoded_steps)
    if same:
        for i in range(0, len(decoded_steps), 8):                 # one gather launch per 8 steps
            ops.bilinear_lift_multi(decoded_steps[i:i + 8], size, out=out, out_channel=i, gray=True)
    else:
        for i, d in enumerate(decoded_steps):
            ops.bilinear_lift(d, size, out=out, out_channel=i, gray=True)
    return out


def feature_concat_autograd(decoded_steps: Sequence[torch.Tensor], size=(64, 64)) -> torch.Tensor:
    """``feature_concat`` for the training caller (ldiffusion.py:240-252): the same kernels in the forward, their
    adjoint in the backward, so the InfoNCE loss on the features reaches the decoder outputs (fp32 [B,n,h,w];
    the concat itself is torch.cat here because autograd needs separate nodes)."""
    return torch.cat([ops.bilinear_lift_autograd(d, size, gray=True) for d in decoded_steps], dim=1)


def label_down(label: torch.Tensor, size=(64, 64)) -> torch.Tensor:
    """ldiffusion.py:224-226: uint8 [B,1,H,W] -> float -> bilinear -> uint8 (truncation)."""
    if label.dtype != torch.uint8:
        label = label.to(torch.uint8)
    return ops.bilinear_lift(label.contiguous(), size)


def rgb_up(rgb: torch.Tensor, size=(1024, 1024)) -> torch.Tensor:
    """ldiffusion.py:251."""
    return ops.bilinear_lift(rgb, size)


def generate_title(n):
    """CSV header of pixel_latent_vector.py:49-56."""
    return ["Pixel No."] + [f"Sample {i + 1}" for i in range(n)] + ["Category"]


def write_pixel_csv(path: str, vectors: torch.Tensor):
    """pixel_latent_vector.py:95-101 for one image: vectors uint8 [H,W,n+1]."""
    v = vectors.cpu().numpy()
    H, W, n1 = v.shape
    with open(path, mode="w", newline="") as f:
        wr = csv.writer(f)
        wr.writerow(generate_title(n1 - 1))
        for i in range(H):
            for j in range(W):
                wr.writerow([(i, j)] + v[i, j].tolist())


def pixel_latent_vector(pipeline, vae, unet, num_inference_steps, train_loader=None,
                        text_embeddings=None, out_dir: Optional[str] = None, write_csv: bool = True):
    """Drop-in for ``pixel_latent_vector.pixel_latent_vector`` (``:58-102``).

    ``pipeline`` needs ``.scheduler`` (diffusers duck-type) and the callables the
    reference uses: ``vae.encode(x).latent_dist.mean``, ``unet(latents, t,
    text_embeddings)[0]`` and ``vae.decode(latents / 0.18215).sample``.  The
    reference reads a module-global ``train_loader``; it is a parameter here.
    Returns the list of per-image PixelVectorBuilder objects.
    """
    if train_loader is None:
        raise ValueError("train_loader is required")
    if out_dir is None:
        out_dir = f"eval/vector_set/{datetime.now().strftime('%y_%m_%d')}"
    if write_csv:
        os.makedirs(out_dir, exist_ok=True)
    results = []
    device = torch.device("cuda", torch.cuda.current_device())
    sched = pipeline.scheduler
    for image_index, (image, label) in enumerate(train_loader):
        image, label = image.to(device), label.to(device)
        with torch.no_grad():
            latents = vae.encode(image).latent_dist.mean
            sched.set_timesteps(num_inference_steps - 1, device=device)
            B, _, H, W = image.shape
            pv = PixelVectorBuilder(B, H, W, num_inference_steps, device)
            for i, t in enumerate(sched.timesteps):
                latents = sched.scale_model_input(latents, t)
                output = unet(latents, t, text_embeddings)
                latents = sched.step(output[0], t, latents).prev_sample
                decoded = vae.decode(latents / 0.18215).sample
                pv.add_step(i, decoded.contiguous())
            pv.set_label((label.reshape(B, H, W) if label.dtype == torch.uint8
                          else label.reshape(B, H, W)).to(torch.uint8))
        if write_csv:
            write_pixel_csv(os.path.join(out_dir, f"pixel_dict_{image_index}.csv"), pv.vectors[0])
        results.append(pv)
    return results
