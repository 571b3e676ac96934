"""Random-init stand-ins for the out-of-scope models, with the interfaces the
reference's hot loops call (diffusers / torchvision / Cellpose are not installed
in the build image and there is no network for weights).

They exist so that the drop-in seams (``Segmentor``, ``LDiffusionModel``,
``pixel_latent_vector``) can be exercised end to end and so that the hot-path
kernels see tensors of the right shape, dtype and layout.  They are plain torch
modules (cuDNN/cuBLAS library calls): NOT part of the measured hot path and not
an implementation of SD-v1.5.  Shapes follow SD-v1.5: VAE 3<->4 channels, x8
spatial, scaling factor 0.18215 applied by the caller; UNet 4->4 channels at
latent resolution, cross-attention width 768.
"""
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .scheduler import LaplacePLMSScheduler


class StandInVAE(nn.Module):
    def __init__(self, width: int = 16):
        super().__init__()
        self.enc = nn.Sequential(nn.Conv2d(3, width, 3, 2, 1), nn.SiLU(), nn.Conv2d(width, width, 3, 2, 1), nn.SiLU(),
                                 nn.Conv2d(width, 4, 3, 2, 1))
        self.dec = nn.Sequential(nn.Conv2d(4, width, 3, 1, 1), nn.SiLU(), nn.Upsample(scale_factor=2),
                                 nn.Conv2d(width, width, 3, 1, 1), nn.SiLU(), nn.Upsample(scale_factor=2),
                                 nn.Conv2d(width, width, 3, 1, 1), nn.SiLU(), nn.Upsample(scale_factor=2),
                                 nn.Conv2d(width, 3, 3, 1, 1))
        self.config = SimpleNamespace(scaling_factor=0.18215)

    def encode(self, x):
        mean = self.enc(x)
        return SimpleNamespace(latent_dist=SimpleNamespace(mean=mean, sample=lambda: mean))

    def decode(self, z):
        return SimpleNamespace(sample=self.dec(z))


class StandInUNet(nn.Module):
    def __init__(self, width: int = 32, cross_attention_dim: int = 768):
        super().__init__()
        self.inp = nn.Conv2d(4, width, 3, 1, 1)
        self.t_proj = nn.Linear(1, width)
        self.c_proj = nn.Linear(cross_attention_dim, width)
        self.mid = nn.Conv2d(width, width, 3, 1, 1)
        self.out = nn.Conv2d(width, 4, 3, 1, 1)
        self.config = SimpleNamespace(cross_attention_dim=cross_attention_dim)

    def forward(self, sample, timestep, encoder_hidden_states=None, **kw):
        t = torch.as_tensor(timestep, dtype=sample.dtype, device=sample.device).reshape(-1, 1) / 1000.0
        h = self.inp(sample) + self.t_proj(t)[:, :, None, None]
        if encoder_hidden_states is not None:
            h = h + self.c_proj(encoder_hidden_states.to(sample.dtype).mean(1))[:, :, None, None]
        h = self.mid(F.silu(h))
        if kw.get("mid_block_additional_residual") is not None:          # ControlNet hook (segmentor.py:366-372)
            h = h + kw["mid_block_additional_residual"]
        h = self.out(F.silu(h))
        return UNetOutput(h)


class StandInControlNet(nn.Module):
    """``ControlNetModel`` as far as segmentor.py:357-363 touches it: (down residuals, mid residual)."""

    def __init__(self, width: int = 32):
        super().__init__()
        self.cond = nn.Conv2d(3, width, 8, 8)
        self.inp = nn.Conv2d(4, width, 3, 1, 1)

    def forward(self, sample, timestep, encoder_hidden_states=None, controlnet_cond=None, return_dict=False):
        mid = F.silu(self.inp(sample) + self.cond(controlnet_cond.to(sample.dtype)))
        return [mid], mid


class UNetOutput(tuple):
    """Supports both ``output[0]`` (segmentor.py:103) and ``output.sample`` (ldiffusion.py:238)."""

    def __new__(cls, sample):
        return super().__new__(cls, (sample,))

    @property
    def sample(self):
        return self[0]


class StandInTextEncoder(nn.Module):
    def __init__(self, hidden: int = 768, vocab: int = 49408):
        super().__init__()
        self.emb = nn.Embedding(vocab, hidden)
        self.config = SimpleNamespace(hidden_size=hidden)

    def forward(self, input_ids):
        return {"last_hidden_state": self.emb(input_ids)}


class StandInTokenizer:
    def __call__(self, prompts, **kw):
        import zlib
        ids = [[49406] + np.random.default_rng(zlib.crc32(str(p).encode())).integers(1000, 2000, 5).tolist() + [49407]
               for p in prompts]                                   # a prompt always maps to the same ids
        return {"input_ids": ids}


class StandInPipeline:
    """What ``StableDiffusionImg2ImgPipeline.from_pretrained`` returns, as far as the
    reference's loops touch it; the scheduler is the real product scheduler."""

    def __init__(self, device="cuda", dtype=torch.float32, seed: int = 0):
        torch.manual_seed(seed)
        self.vae = StandInVAE().to(device, dtype).eval()
        self.unet = StandInUNet().to(device, dtype).eval()
        self.text_encoder = StandInTextEncoder().to(device, dtype).eval()
        self.tokenizer = StandInTokenizer()
        self.scheduler = LaplacePLMSScheduler()
        self.device = torch.device(device)

    def to(self, device):
        return self


class StandInCellModel(nn.Module):
    """CellSegClassifier's out-of-scope front half (Cellpose instances + ResNet-152
    features, conductor.py:175-216) replaced by a grid of square 'cells' and random
    features; the in-scope tail (Linear(256,K) head + painting) is the product's."""

    def __init__(self, num_classes: int, cell: int = 24, device="cuda"):
        super().__init__()
        self.num_classes, self.cell = num_classes, cell
        self.classifier = nn.Linear(256, num_classes).to(device)

    @torch.no_grad()
    def instances(self, image_np):
        H, W = image_np.shape[:2]
        dev = self.classifier.weight.device
        ys = torch.arange(H, device=dev) // (2 * self.cell)
        xs = torch.arange(W, device=dev) // (2 * self.cell)
        ncol = (W + 2 * self.cell - 1) // (2 * self.cell)
        inst = (ys[:, None] * ncol + xs[None, :] + 1).to(torch.int32)
        gap = ((torch.arange(H, device=dev) % (2 * self.cell)) >= self.cell)[:, None] | \
              ((torch.arange(W, device=dev) % (2 * self.cell)) >= self.cell)[None, :]
        inst[gap] = 0
        ids = torch.unique(inst)
        ids = ids[ids != 0].to(torch.int32)
        g = torch.Generator(device="cpu").manual_seed(int(ids.numel()))
        feats = torch.randn(ids.numel(), 256, generator=g).to(dev)
        return inst.contiguous(), feats, ids


class StandInTissueModel(nn.Module):
    """TissueSegNet's out-of-scope body (ConvNeXt/CBAM/ASPP/3x3 conv, conductor.py:114-126)
    replaced by a strided conv; ``head`` is the real final 1x1 conv (conductor.py:127)."""

    def __init__(self, num_classes: int, device="cuda"):
        super().__init__()
        self.body = nn.Sequential(nn.Conv2d(3, 64, 7, 4, 3), nn.ReLU(), nn.Conv2d(64, 256, 3, 8, 1), nn.ReLU()).to(device)
        self.head = nn.Conv2d(256, num_classes, 1).to(device)
        self.num_classes = num_classes

    @torch.no_grad()
    def features(self, x):
        return self.body(x)
