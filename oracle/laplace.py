"""Oracle for stage a-1: Laplace forward noising (q_sample).

Test infrastructure only (see ``oracle/__init__.py``).

Reference: ``ldiffusion.py:233-237``::

    scale = sqrt(1 - alphas_cumprod[t])
    noise = torch.distributions.Laplace(0, scale).sample(latents.shape)
    noisy = latents + noise

and the multimodal variant ``segmentor.py:344-345`` (``latents + Laplace(0,1) *
depth``).  The sampler is torch's own ``Laplace.rsample`` (third-party to the
reference, importable here): ``u ~ U(eps - 1, 1)``, ``noise = loc - scale *
sign(u) * log1p(-|u|)``.  Note there is no sqrt(alpha_bar) scaling of x.
"""
import numpy as np
import torch

from .scheduler import PNDMOracle


def laplace_scale(t, scheduler=None):
    """ldiffusion.py:234: b_t = sqrt(1 - alpha_bar_t) (fp32)."""
    s = scheduler or PNDMOracle()
    return torch.sqrt(1 - s.alphas_cumprod[int(t)])


def laplace_from_uniform_chain(u, scale):
    """The transform inside torch.distributions.Laplace.rsample, op for op:
    ``loc - scale * u.sign() * torch.log1p(-u.abs())`` with loc = 0."""
    scale = torch.as_tensor(scale, dtype=torch.float32)
    loc = torch.zeros_like(scale)
    return loc - scale * u.sign() * torch.log1p(-u.abs())


def qsample_chain(latents, t, generator=None, scheduler=None):
    """ldiffusion.py:233-237 with torch's sampler.  Returns (noisy, noise)."""
    b = laplace_scale(t, scheduler)
    if generator is None:
        noise = torch.distributions.Laplace(0, b).sample(latents.shape).to(torch.float32)
    else:
        finfo = torch.finfo(torch.float32)
        u = torch.empty(latents.shape, dtype=torch.float32).uniform_(finfo.eps - 1, 1, generator=generator)
        noise = laplace_from_uniform_chain(u, b)
    return (latents + noise).to(torch.float32), noise


def qsample_injected(latents, noise):
    """ldiffusion.py:237 with an injected noise tensor: one fp32 add."""
    return (latents + noise).to(torch.float32)


def depth_noising_chain(latents, depth_resized, noise=None, generator=None, x_mul=None):
    """segmentor.py:339-345 (``ldiffusion_augment_for_multimodal``), op for op::

        latents = vae.encode(rgb_i).latent_dist.sample() * 0.18215      # x_mul, when given
        depth_resized = depth_resized.repeat(1, latents.shape[1], 1, 1)
        noise = torch.distributions.laplace.Laplace(0.0, 1.0).sample(latents.shape)
        latents_noisy = latents + noise * depth_resized

    ``depth_resized`` is [B,1,h,w] (repeated here as the reference does) or already [B,C,h,w].
    ``noise`` injects the Laplace(0,1) tensor; otherwise torch's sampler draws it (``generator``
    replays ``rsample`` on a seeded uniform).  Returns (latents_noisy, noise)."""
    if x_mul is not None:
        latents = latents * x_mul
    if depth_resized.shape[1] != latents.shape[1]:
        depth_resized = depth_resized.repeat(1, latents.shape[1], 1, 1)
    if noise is None:
        if generator is None:
            noise = torch.distributions.laplace.Laplace(0.0, 1.0).sample(latents.shape)
        else:
            finfo = torch.finfo(torch.float32)
            u = torch.empty(latents.shape, dtype=torch.float32).uniform_(finfo.eps - 1, 1, generator=generator)
            noise = laplace_from_uniform_chain(u, 1.0)
    return latents + noise * depth_resized, noise


def depth_unnoise_chain(latents_noisy, noise_pred, depth_resized, out_div=None):
    """segmentor.py:375,379: ``latents_denoised = latents_noisy - noise_pred * depth_resized`` and, when
    ``out_div`` is given, the ``latents_denoised / 0.18215`` handed to ``vae.decode``."""
    if depth_resized.shape[1] != latents_noisy.shape[1]:
        depth_resized = depth_resized.repeat(1, latents_noisy.shape[1], 1, 1)
    out = latents_noisy - noise_pred * depth_resized
    return out if out_div is None else out / out_div


# ---- the counter-based generator of the CUDA kernel, restated -------------
# (Philox4x32-10 is a published algorithm: Salmon et al., SC'11.  torch's RNG
# stream cannot and need not be matched; the kernel's own stream is pinned by
# this restatement so that its noise is reproducible on the CPU.)

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)


def philox4x32_10(counter, key):
    """counter: uint32 [N,4]; key: uint32 [2] -> uint32 [N,4]."""
    c = np.array(counter, dtype=np.uint32, copy=True)
    k0, k1 = np.uint32(key[0]), np.uint32(key[1])
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _M0 * c[:, 0].astype(np.uint64)
            p1 = _M1 * c[:, 2].astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
            c = np.stack([hi1 ^ c[:, 1] ^ k0, lo1, hi0 ^ c[:, 3] ^ k1, lo0], axis=1)
            k0 = np.uint32(k0 + _W0)
            k1 = np.uint32(k1 + _W1)
    return c


def philox_uniform_pm1(n, seed, offset):
    """The kernel's uniform stream: element i uses word (i % 4) of
    philox(counter = offset + i // 4, key = seed); with k = word >> 9 (23 bits),
    u = (2k + 1 - 2^23) * 2^-23: odd multiples of 2^-23, symmetric, never 0 and
    strictly inside (-1, 1); every step is exact in fp32.  Returns fp32 [n]."""
    nblk = (n + 3) // 4
    ctr = np.uint64(offset) + np.arange(nblk, dtype=np.uint64)
    counter = np.zeros((nblk, 4), dtype=np.uint32)
    counter[:, 0] = ctr.astype(np.uint32)
    counter[:, 1] = (ctr >> np.uint64(32)).astype(np.uint32)
    key = np.array([np.uint64(seed) & np.uint64(0xFFFFFFFF), np.uint64(seed) >> np.uint64(32)],
                   dtype=np.uint64).astype(np.uint32)
    words = philox4x32_10(counter, key).reshape(-1)[:n]
    k = (words >> np.uint32(9)).astype(np.int64)
    return (2 * k + 1 - (1 << 23)).astype(np.float32) * np.float32(2.0 ** -23)


def laplace_philox(n, scale, seed, offset):
    """noise the CUDA kernel draws in Philox mode (transform as in rsample)."""
    u = torch.from_numpy(philox_uniform_pm1(n, seed, offset))
    return laplace_from_uniform_chain(u, scale)
