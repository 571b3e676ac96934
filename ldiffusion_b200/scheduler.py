"""Laplace forward noising + PLMS reverse scheduler with the diffusers method set.

``LaplacePLMSScheduler`` is a drop-in for ``pipeline.scheduler`` at the
reference's call sites (``ldiffusion.py:198,229-237``; ``segmentor.py:100-104``,
``:438-445``, ``:520-527``; ``utils.py:196-202``; ``pixel_latent_vector.py:74-79``;
``sample.py:57-64``): ``set_timesteps``, ``timesteps``, ``scale_model_input``,
``step(...).prev_sample``, ``alphas_cumprod``, ``init_noise_sigma``.  It is
configured as SD-v1.5's PNDMScheduler (skip_prk_steps, steps_offset 1, leading
spacing, epsilon prediction).

The bookkeeping (which Adams-Bashforth order applies, the duplicated second
timestep, the stashed first sample) and the three per-step scalars are host
work; the tensor update is ONE fused kernel launch per step
(``ldiff_plms_step``) instead of the 12-15 eager launches of the reference, and
the alpha-bar table stays on the host so no step synchronises.  The scalars are
computed with the same 0-dim fp32 tensor arithmetic the reference performs, so
the kernel's result is bit-identical to the eager chain in fp32.
"""
from dataclasses import dataclass

import numpy as np
import torch

from . import ops


@dataclass
class SchedulerOutput:
    prev_sample: torch.Tensor


class LaplacePLMSScheduler:
    order = 1

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.00085,
                 beta_end: float = 0.012, steps_offset: int = 1, set_alpha_to_one: bool = False):
        self.num_train_timesteps = num_train_timesteps
        self.steps_offset = steps_offset
        # scaled_linear schedule, fp32 host table (never moved to the device: indexing it
        # with a host int costs no synchronisation)
        self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps,
                                    dtype=torch.float32) ** 2
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.pndm_order = 4
        self.num_inference_steps = None
        self.timesteps = None
        self._host_timesteps = None
        self.ets = []
        self.counter = 0
        self.cur_sample = None

    # -- diffusers surface ---------------------------------------------------
    def set_timesteps(self, num_inference_steps: int, device=None):
        if num_inference_steps < 1:
            raise ValueError("num_inference_steps must be >= 1")
        self.num_inference_steps = num_inference_steps
        ratio = self.num_train_timesteps // num_inference_steps
        base = (np.arange(0, num_inference_steps) * ratio).round() + self.steps_offset
        plms = np.concatenate([base[:-1], base[-2:-1], base[-1:]])[::-1].copy().astype(np.int64)
        self._host_timesteps = [int(t) for t in plms]
        self.timesteps = torch.from_numpy(plms).to(device) if device is not None else torch.from_numpy(plms)
        self.ets = []
        self.counter = 0
        self.cur_sample = None

    def scale_model_input(self, sample, timestep=None):
        return sample

    def _as_int(self, timestep):
        """Timesteps iterated from ``self.timesteps`` may be device scalars.  An element (or a slice's
        element) of ``self.timesteps`` is resolved from the host copy BY ITS POSITION IN THAT TENSOR —
        the value the caller passed, without a synchronisation, also for ``timesteps[t_start:]`` or a
        reordered walk; any other device tensor is read back (one sync, as diffusers does)."""
        if isinstance(timestep, torch.Tensor):
            ts = self.timesteps
            if (timestep.is_cuda and timestep.dim() == 0 and ts is not None and ts.is_cuda
                    and self._host_timesteps is not None
                    and timestep.untyped_storage().data_ptr() == ts.untyped_storage().data_ptr()):
                idx = timestep.storage_offset() - ts.storage_offset()
                if 0 <= idx < len(self._host_timesteps):
                    return self._host_timesteps[idx]
            return int(timestep)
        return int(timestep)

    def plan_step(self, timestep: int):
        """Host part of one step: returns (mode, uses_stashed_sample, sample_coeff,
        alpha_diff, denom) for the call about to happen and does not touch tensors."""
        ratio = self.num_train_timesteps // self.num_inference_steps
        prev_timestep = timestep - ratio
        n_hist = len(self.ets)
        if self.counter != 1:
            n_hist = min(n_hist, 3) + 1
        else:
            prev_timestep = timestep
            timestep = timestep + ratio
        if n_hist == 1 and self.counter == 0:
            mode = 0
        elif n_hist == 1 and self.counter == 1:
            mode = 1
        elif n_hist == 2:
            mode = 2
        elif n_hist == 3:
            mode = 3
        else:
            mode = 4
        a_t = self.alphas_cumprod[timestep]
        a_p = self.alphas_cumprod[prev_timestep] if prev_timestep >= 0 else self.final_alpha_cumprod
        b_t = 1 - a_t
        b_p = 1 - a_p
        sample_coeff = (a_p / a_t) ** 0.5
        denom = a_t * b_p ** 0.5 + (a_t * b_t * a_p) ** 0.5
        return mode, sample_coeff.item(), (a_p - a_t).item(), denom.item()

    def step(self, model_output, timestep, sample, return_dict: bool = True, out=None, noise_of=None):
        if self.num_inference_steps is None:
            raise ValueError(
                "Number of inference steps is 'None', you need to run 'set_timesteps' after creating the scheduler")
        t = self._as_int(timestep)
        mode, sc, dA, denom = self.plan_step(t)
        if self.counter != 1:
            self.ets = self.ets[-3:]
            self.ets.append(model_output)
        if mode == 0:
            self.cur_sample = sample
            eps = [model_output]
        elif mode == 1:
            eps = [model_output, self.ets[-1]]
            sample = self.cur_sample
            self.cur_sample = None
        else:
            eps = self.ets[::-1]
        if noise_of is None:
            prev = ops.plms_step(sample, eps, mode, sc, dA, denom, out=out)
        else:
            prev, noisy = ops.plms_step_noise(sample, eps, mode, sc, dA, denom, noise_of["clean"],
                                              self.laplace_scale(noise_of.get("timestep", t)),
                                              noise=noise_of.get("noise"), u=noise_of.get("u"),
                                              seed=noise_of.get("seed", 0), offset=noise_of.get("offset", 0),
                                              out=out, noisy_out=noise_of.get("out"))
        self.counter += 1
        if noise_of is not None:
            return SchedulerOutput(prev_sample=prev), noisy
        if not return_dict:
            return (prev,)
        return SchedulerOutput(prev_sample=prev)

    def step_then_noise(self, model_output, timestep, sample, clean, *, noise_timestep=None, noise=None, u=None,
                        seed: int = 0, offset: int = 0, out=None, noisy_out=None):
        """The reverse update of ``step`` AND the Laplace forward noising of ``clean`` at
        ``noise_timestep`` (default: ``timestep``) in ONE kernel launch — the two latent-sized ops the
        reference runs per loop iteration (segmentor.py:100-104; ldiffusion.py:233-237).  Returns
        (prev_sample, noisy); bit-identical to ``step`` + ``add_laplace_noise``."""
        t = self._as_int(timestep)
        spec = {"clean": clean, "timestep": t if noise_timestep is None else int(noise_timestep), "noise": noise,
                "u": u, "seed": seed, "offset": offset, "out": noisy_out}
        res, noisy = self.step(model_output, timestep, sample, out=out, noise_of=spec)
        return res.prev_sample, noisy

    def __len__(self):
        return self.num_train_timesteps

    # -- Laplace forward noising (ldiffusion.py:233-237) ----------------------
    def laplace_scale(self, timestep) -> float:
        """b_t = sqrt(1 - alpha_bar_t), fp32, from the host table."""
        return torch.sqrt(1 - self.alphas_cumprod[int(timestep)]).item()

    def add_laplace_noise(self, latents, timestep, *, noise=None, u=None, seed: int = 0, offset: int = 0,
                          return_noise: bool = False):
        """noisy = latents + Laplace(0, b_t) in one fused launch (Philox sampling, the
        inverse-CDF transform and the add never leave registers)."""
        return ops.laplace_qsample(latents, self.laplace_scale(timestep), noise=noise, u=u, seed=seed,
                                   offset=offset, return_noise=return_noise)
