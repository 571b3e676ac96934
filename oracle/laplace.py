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


def philox4x32(counter, key, rounds=10):
    """counter: uint32 [N,4]; key: uint32 [2] -> uint32 [N,4] (Philox4x32-R, Salmon et al. SC'11)."""
    c = np.array(counter, dtype=np.uint32, copy=True)
    k0, k1 = np.uint32(key[0]), np.uint32(key[1])
    with np.errstate(over="ignore"):
        for _ in range(rounds):
            p0 = _M0 * c[:, 0].astype(np.uint64)
            p1 = _M1 * c[:, 2].astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
            c = np.stack([hi1 ^ c[:, 1] ^ k0, lo1, hi0 ^ c[:, 3] ^ k1, lo0], axis=1)
            k0 = np.uint32(k0 + _W0)
            k1 = np.uint32(k1 + _W1)
    return c


def philox4x32_10(counter, key):
    return philox4x32(counter, key, 10)


def philox_words(n, seed, offset, rounds=10):
    """The kernel's word stream: element i uses word (i % 4) of philox(counter = offset + i // 4, key = seed)."""
    nblk = (n + 3) // 4
    ctr = np.uint64(offset) + np.arange(nblk, dtype=np.uint64)
    counter = np.zeros((nblk, 4), dtype=np.uint32)
    counter[:, 0] = ctr.astype(np.uint32)
    counter[:, 1] = (ctr >> np.uint64(32)).astype(np.uint32)
    key = np.array([np.uint64(seed) & np.uint64(0xFFFFFFFF), np.uint64(seed) >> np.uint64(32)],
                   dtype=np.uint64).astype(np.uint32)
    return philox4x32(counter, key, rounds).reshape(-1)[:n]


def philox_uniform_pm1(n, seed, offset, rounds=10):
    """The kernel's uniform stream: bits [22:0] of a word are the magnitude m (23 bits), bit 31 the sign:
    u = +- m * 2^-23 — a symmetric 24-bit uniform strictly inside (-1, 1) (u = 0 with probability 2^-23); every
    step is exact in fp32.  Returns fp32 [n] (a zero keeps its sign bit: -0.0 where bit 31 is set)."""
    words = philox_words(n, seed, offset, rounds)
    mag = (words & np.uint32(0x007FFFFF)).astype(np.float32) * np.float32(2.0 ** -23)
    return np.where(words >> np.uint32(31), -mag, mag).astype(np.float32)


def laplace_philox(n, scale, seed, offset, rounds=10, storage="f32"):
    """noise the CUDA kernel draws in Philox mode.

    ``storage="f32"``: one 32-bit word per element (24-bit symmetric uniform), torch's rsample transform
    -b * sign(u) * log1p(-|u|) (the noise carries the sign of u, log1p(-|u|) being negative).
    ``storage="bf16"``: TWO variates per word (see ``philox_half_stream``): 15-bit magnitude + sign, the top
    magnitude cell refined by a second 23-bit draw (the exponential tail is memoryless), so
    |noise| = b * (15 ln 2 - ln(1 - m2 * 2^-23)) there.  ``rounds``: 10 by default (LDIFF_TUNE_PHILOX_ROUNDS = 7
    selects the 7-round stream)."""
    if storage == "bf16":
        lg2, neg = philox_half_stream(n, seed, offset, rounds)
        mag = torch.from_numpy(lg2.astype(np.float64)) * (-float(np.log(2.0))) * float(scale)
        return (torch.where(torch.from_numpy(neg), -mag, mag)).to(torch.float32)
    u = torch.from_numpy(philox_uniform_pm1(n, seed, offset, rounds))
    return laplace_from_uniform_chain(u, scale)


def philox_half_stream(n, seed, offset, rounds=10):
    """The bf16-storage stream: element i uses half (i % 2) of word ((i % 8) // 2) of philox(counter = offset + i // 8,
    key = seed); half 0 = bits [31:16], half 1 = bits [15:0]; in a half, bit 15 is the sign and bits [14:0] the
    magnitude m: |u| = m * 2^-15.  m = 2^15 - 1 (the top cell) is refined: word 0 of philox(counter with its third
    word = 1 + (i % 8)) gives m2 = its low 23 bits and log2(1 - |u|) := -15 + log2(1 - m2 * 2^-23).
    Returns (log2(1 - |u|) as float64 [n], sign-is-negative bool [n])."""
    ngrp = (n + 7) // 8
    ctr = np.uint64(offset) + np.arange(ngrp, dtype=np.uint64)
    counter = np.zeros((ngrp, 4), dtype=np.uint32)
    counter[:, 0] = ctr.astype(np.uint32)
    counter[:, 1] = (ctr >> np.uint64(32)).astype(np.uint32)
    key = np.array([np.uint64(seed) & np.uint64(0xFFFFFFFF), np.uint64(seed) >> np.uint64(32)],
                   dtype=np.uint64).astype(np.uint32)
    words = philox4x32(counter, key, rounds)                                   # [ngrp, 4]
    halves = np.stack([words >> np.uint32(16), words & np.uint32(0xFFFF)], axis=2).reshape(ngrp, 8)
    halves = halves.reshape(-1)[:n].astype(np.uint32)
    m = (halves & np.uint32(0x7FFF)).astype(np.float64)
    neg = (halves >> np.uint32(15)).astype(bool)
    lg2 = np.log2(1.0 - m * 2.0 ** -15)
    top = np.nonzero((halves & np.uint32(0x7FFF)) == 0x7FFF)[0]
    for i in top:                                                              # rare: 2^-15 of the elements
        c = counter[i // 8].copy()
        c[2] = np.uint32(1 + (i % 8))
        m2 = float(philox4x32(c[None, :], key, rounds)[0, 0] & np.uint32(0x007FFFFF))
        lg2[i] = -15.0 + np.log2(1.0 - m2 * 2.0 ** -23)
    return lg2, neg
