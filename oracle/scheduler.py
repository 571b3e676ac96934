"""Oracle for stage a-2: the few-step reverse scheduler update.

Test infrastructure only (see ``oracle/__init__.py``).  **Parity unpinned**:
the algorithm lives in ``diffusers==0.34.0`` (``environment.yml:42``), class
``PNDMScheduler`` (SD-v1.5's default scheduler), which is not vendored in the
reference and not installed in the build image.  This restates the published
algorithm with the SD-v1.5 ``scheduler_config.json`` values::

    beta_start 0.00085, beta_end 0.012, beta_schedule scaled_linear,
    num_train_timesteps 1000, skip_prk_steps true, set_alpha_to_one false,
    steps_offset 1, timestep_spacing leading, prediction_type epsilon

Anchors: the reference's call sites (``segmentor.py:100-104``, ``:438-445``,
``:520-527``; ``utils.py:196-202``; ``pixel_latent_vector.py:74-79``;
``sample.py:57-64``; ``ldiffusion.py:198,229-234``) and the known-answer
constants of SURVEY.md §8(a-2), checked in ``tests/test_oracle_ops.py`` together with two
first-principles pins (the published transfer formula in fp64; exactness of the whole loop,
quirks included, for a perfect noise prediction).

All tensor arithmetic is torch-CPU fp32, op by op, in the published order, so
the CUDA ``plms_step`` kernel (which uses round-to-nearest intrinsics in the
same order) can be compared bit for bit.
"""
import numpy as np
import torch


class PNDMOracle:
    order = 1

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012,
                 steps_offset=1, set_alpha_to_one=False):
        self.num_train_timesteps = num_train_timesteps
        self.steps_offset = steps_offset
        self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps,
                                    dtype=torch.float32) ** 2
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.pndm_order = 4
        self.num_inference_steps = None
        self.timesteps = None
        self.ets = []
        self.counter = 0
        self.cur_sample = None

    def set_timesteps(self, num_inference_steps, device=None):
        self.num_inference_steps = num_inference_steps
        ratio = self.num_train_timesteps // num_inference_steps
        base = (np.arange(0, num_inference_steps) * ratio).round() + self.steps_offset
        # skip_prk_steps: the second-highest timestep is visited twice
        plms = np.concatenate([base[:-1], base[-2:-1], base[-1:]])[::-1].copy()
        self.timesteps = torch.from_numpy(plms.astype(np.int64))
        self.ets = []
        self.counter = 0
        self.cur_sample = None

    def scale_model_input(self, sample, timestep=None):
        return sample

    def step(self, model_output, timestep, sample):
        """Returns prev_sample (the reference reads ``.prev_sample``)."""
        if self.num_inference_steps is None:
            raise ValueError("call set_timesteps first")
        timestep = int(timestep)
        ratio = self.num_train_timesteps // self.num_inference_steps
        prev_timestep = timestep - ratio
        if self.counter != 1:
            self.ets = self.ets[-3:]
            self.ets.append(model_output)
        else:
            prev_timestep = timestep
            timestep = timestep + ratio

        if len(self.ets) == 1 and self.counter == 0:
            self.cur_sample = sample
        elif len(self.ets) == 1 and self.counter == 1:
            model_output = (model_output + self.ets[-1]) / 2
            sample = self.cur_sample
            self.cur_sample = None
        elif len(self.ets) == 2:
            model_output = (3 * self.ets[-1] - self.ets[-2]) / 2
        elif len(self.ets) == 3:
            model_output = (23 * self.ets[-1] - 16 * self.ets[-2] + 5 * self.ets[-3]) / 12
        else:
            model_output = (1 / 24) * (55 * self.ets[-1] - 59 * self.ets[-2]
                                       + 37 * self.ets[-3] - 9 * self.ets[-4])

        prev = self._get_prev_sample(sample, timestep, prev_timestep, model_output)
        self.counter += 1
        return prev

    def _get_prev_sample(self, sample, timestep, prev_timestep, model_output):
        a_t = self.alphas_cumprod[timestep]
        a_p = self.alphas_cumprod[prev_timestep] if prev_timestep >= 0 else self.final_alpha_cumprod
        b_t = 1 - a_t
        b_p = 1 - a_p
        sample_coeff = (a_p / a_t) ** 0.5
        denom = a_t * b_p ** 0.5 + (a_t * b_t * a_p) ** 0.5
        return sample_coeff * sample - (a_p - a_t) * model_output / denom


def sample_loop(x0, eps_list, num_set_timesteps):
    """The reference's sampling loop with the UNet replaced by given outputs
    (segmentor.py:100-104 et al.).  Returns the list of latents after each step."""
    s = PNDMOracle()
    s.set_timesteps(num_set_timesteps)
    assert len(eps_list) == len(s.timesteps)
    lat = x0
    out = []
    for eps, t in zip(eps_list, s.timesteps):
        lat = s.scale_model_input(lat, t)
        lat = s.step(eps, t, lat)
        out.append(lat)
    return out
