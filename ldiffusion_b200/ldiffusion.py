"""``LDiffusionModel`` with the reference's interface (``ldiffusion.py:31-324``).

In scope: ``inference`` (dispatch to the Segmentor hot loops) and the tensor part of
one warm-up step (``ldiffusion.py:224-251``: label down-sample, Laplace noising per
timestep, decode -> bilinear -> gray -> concat, final RGB up-sample), exposed as
``laplace_feature_step``.  Out of scope: the DeepSpeed ZeRO-3 engine, AdamW, the
InfoNCE/VGG loss, checkpointing — ``train`` therefore needs an injected
``train_step`` callable and otherwise raises.
"""
import os

import torch

from . import features
from .segmentor import Segmentor


class LDiffusionModel:
    def __init__(self, diffusion_path, level, local_rank=-1, pipeline_loader=None, model_factory=None):
        self.local_rank = local_rank
        self.world_size = int(os.environ.get("WORLD_SIZE", "1"))
        self.global_rank = int(os.environ.get("RANK", "0"))
        self.is_distributed = self.world_size > 1
        if not torch.cuda.is_available():
            raise RuntimeError("ldiffusion_b200.LDiffusionModel needs a CUDA device (no CPU fallback)")
        if self.local_rank is None or self.local_rank < 0:
            self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.device = torch.device(f"cuda:{self.local_rank}")
        self.level = level
        self.diffusion_path = diffusion_path
        self.pipeline, self.vae = None, None
        self.linear_layer = None
        self._pipeline_loader, self._model_factory = pipeline_loader, model_factory

    def _is_main_process(self):
        return self.global_rank == 0

    def _segmentor(self, train_loader, val_loader, num_classes):
        return Segmentor(train_loader, val_loader, self.level, num_classes,
                         pipeline_loader=self._pipeline_loader, model_factory=self._model_factory)

    def inference(self, image_path, ldiffusion_weight, segmentor_weight, num_classes):
        """ldiffusion.py:317-324 -> (PIL.Image, np.uint8 mask [H0,W0])."""
        seg = self._segmentor(None, None, num_classes)
        if self.level == "tissue":
            return seg.inference_tissue_model_nnUNetv2(image_path, self.diffusion_path, ldiffusion_weight,
                                                       segmentor_weight)
        elif self.level == "cell":
            return seg.inference_cell_model(image_path, self.diffusion_path, ldiffusion_weight, segmentor_weight)
        raise ValueError("Invalid level specified. Choose 'tissue' or 'cell'.")

    @torch.no_grad()
    def laplace_feature_step(self, latents, label, scheduler, unet, vae, text_embeddings, num_inference_steps,
                             seed: int = 0, noise=None):
        """The tensor work of one warm-up batch, ldiffusion.py:224-251.

        latents [B,4,h,w] clean VAE means; label uint8 [B,1,H,W].  Returns
        (decoded_image_rgb [B,3,1024,1024], decoded_image_gray [B,n,64,64], label64 uint8 [B,1,64,64]).
        Note the reference never updates ``latents`` in this loop."""
        label64 = features.label_down(label, (64, 64))                               # :224-226
        scheduler.set_timesteps(num_inference_steps, device=latents.device)          # :229
        blocks = (latents.numel() + 3) // 4
        steps, last = [], None
        for i, t in enumerate(scheduler.timesteps):
            x = scheduler.scale_model_input(latents, t)                              # :233
            noisy = scheduler.add_laplace_noise(x, scheduler._host_timesteps[i], seed=seed, offset=i * blocks,
                                                noise=None if noise is None else noise[i])   # :234-237
            denoised = unet(noisy, t, text_embeddings).sample                        # :238 (out of scope)
            last = vae.decode(denoised).sample.contiguous()                          # :240 (out of scope)
            steps.append(last)
        gray = features.feature_concat(steps, (64, 64))                              # :240-247
        rgb64 = features.ops.bilinear_lift(last, (64, 64))                           # :240 (the RGB that is kept)
        rgb = features.rgb_up(rgb64, (1024, 1024))                                   # :251
        return rgb, gray, label64

    def train(self, args, component="all", ldiffusion_weight=None, train_step=None):
        """ldiffusion.py:297-315.  Training orchestration (DeepSpeed ZeRO-3, losses,
        checkpoints) is out of scope; pass ``train_step`` to drive it yourself."""
        if self.level not in ("tissue", "cell"):
            raise ValueError("Invalid level specified. Choose 'tissue' or 'cell'.")
        if train_step is None:
            raise NotImplementedError(
                "training orchestration is outside the scope of ldiffusion_b200 (see DESIGN.md section 7); "
                "the in-scope tensor work of a warm-up step is laplace_feature_step()")
        return train_step(self, args, component, ldiffusion_weight)
