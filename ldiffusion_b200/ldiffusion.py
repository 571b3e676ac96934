"""``LDiffusionModel`` with the reference's interface (``ldiffusion.py:31-324``).

In scope: ``inference`` (dispatch to the Segmentor hot loops) and the tensor part of
one warm-up step (``ldiffusion.py:224-251``: label down-sample, Laplace noising per
timestep, decode -> bilinear -> gray -> concat, final RGB up-sample), exposed as
``laplace_feature_step``; and, as an unmeasured round-2 widening, the same step with
gradients (``warmup_step`` / ``train_ldiffusion``: lift adjoint + InfoNCE kernels, plain
AdamW).  Out of scope: the DeepSpeed ZeRO-3 engine, the VGG content loss, checkpointing,
data loading — ``train`` therefore needs an injected ``train_step`` callable and
otherwise raises.
"""
import os

import torch

from . import features
from .segmentor import Segmentor


class LDiffusionModel:
    def __init__(self, diffusion_path, level, local_rank=-1, pipeline_loader=None, model_factory=None,
                 allow_standins: bool = False):
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
        self._allow_standins = bool(allow_standins)

    def _is_main_process(self):
        return self.global_rank == 0

    def _segmentor(self, train_loader, val_loader, num_classes):
        return Segmentor(train_loader, val_loader, self.level, num_classes,
                         pipeline_loader=self._pipeline_loader, model_factory=self._model_factory,
                         allow_standins=self._allow_standins)

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

    def warmup_step(self, image, label, pipeline, unet, vae, proj, optimizer, num_inference_steps,
                    seed: int = 0, step_index: int = 0, size=(64, 64), noise=None, pairs=None):
        """One batch of the Laplace warm-up, ldiffusion.py:209-255, WITH gradients, on the product kernels
        (round-2 widening; unmeasured).  Per timestep: Laplace noising of the clean latents (one launch, no
        grad), UNet, VAE decode (both out of scope, with grad), bilinear -> gray feature (lift kernel forward,
        adjoint kernel backward); then the pixel-contrastive InfoNCE term on the concatenated features
        (``loss.pixel_contrastive_loss``; the VGG content term of ``loss.py:19-42`` is out of scope), backward,
        gradient clipping at 1.0 and the optimizer step (plain AdamW in place of the DeepSpeed engine,
        ldiffusion.py:165-193).  image [B,3,h,w] float, label uint8 [B,1,H,W].  Returns the detached loss."""
        from . import ops
        from .loss import pixel_contrastive_loss
        dev = self.device
        image = image.to(dev, torch.float32)
        ids = torch.tensor(pipeline.tokenizer(["A pathological slide"] * image.shape[0])["input_ids"],
                           dtype=torch.long, device=dev)                                  # :211-214
        with torch.no_grad():
            text = pipeline.text_encoder(ids)["last_hidden_state"].to(torch.float32)      # :215-216
            label64 = features.label_down(label.to(dev), size)                            # :223-226
            latents = vae.encode(image).latent_dist.mean.to(torch.float32).contiguous()   # :227-228
        text = proj(text)                                                                 # :219
        sched = pipeline.scheduler
        sched.set_timesteps(num_inference_steps, device=dev)                              # :229
        # per-rank, per-purpose Philox keys: every rank draws its own noise for its own batch (the reference's
        # per-process torch generators), and the contrastive sampler never shares counters with the noise stream
        rank = int(getattr(self, "global_rank", 0))
        noise_seed = (int(seed) ^ (rank << 32) ^ 0x4C61706C61636500) & (2 ** 64 - 1)      # "Laplace"
        pair_seed = (int(seed) ^ (rank << 32) ^ 0x496E666F4E434500) & (2 ** 64 - 1)       # "InfoNCE"
        blocks = (latents.numel() + 3) // 4
        n = len(sched.timesteps)
        grays = []
        for i, t in enumerate(sched.timesteps):
            with torch.no_grad():
                x = sched.scale_model_input(latents, t)                                   # :233
                noisy = sched.add_laplace_noise(x, sched._host_timesteps[i], seed=noise_seed,
                                                offset=(step_index * n + i) * blocks,
                                                noise=None if noise is None else noise[i])   # :234-237
            denoised = unet(noisy, t, text).sample                                        # :238
            decoded = vae.decode(denoised.to(torch.float32)).sample                       # :240
            grays.append(ops.bilinear_lift_autograd(decoded, size, gray=True))            # :240-242
        gray = torch.cat(grays, dim=1)                                                    # :244-247
        loss = pixel_contrastive_loss(gray, label64, pairs=pairs, seed=pair_seed, offset=step_index)   # :252
        optimizer.zero_grad(set_to_none=True)
        loss.backward()                                                                   # :254
        params = [p for g in optimizer.param_groups for p in g["params"] if p.grad is not None]
        if params:
            torch.nn.utils.clip_grad_norm_(params, 1.0)                                   # "gradient_clipping": 1.0
        optimizer.step()                                                                  # :255
        return loss.detach()

    def train_ldiffusion(self, args, train_loader, val_loader=None, pipeline=None, log=None):
        """ldiffusion.py:121-295 without DeepSpeed: AdamW(lr 1e-5, betas (0.9, 0.999), eps 1e-8, weight decay
        0.01) over the UNet and the text projection, ``warmup_step`` per batch, the epoch-mean loss (summed
        over ranks when torch.distributed is initialised, :56-64) appended to ``log`` and returned as a list.
        ``train_loader`` yields (image, _, label) like the reference's; checkpoint writing is out of scope."""
        import torch.distributed as dist
        if pipeline is None:
            from .standin import StandInPipeline
            pipeline = StandInPipeline(self.device) if self._pipeline_loader is None else \
                self._pipeline_loader(None, getattr(args, "diffusion_path", self.diffusion_path))[0]
        self.pipeline, self.vae = pipeline, pipeline.vae
        unet = pipeline.unet
        hid, cad = pipeline.text_encoder.config.hidden_size, unet.config.cross_attention_dim
        if self.linear_layer is None or self.linear_layer.in_features != hid or self.linear_layer.out_features != cad:
            self.linear_layer = torch.nn.Linear(hid, cad)                                 # :141-150
        self.linear_layer = self.linear_layer.to(self.device, torch.float32)
        opt = torch.optim.AdamW(list(unet.parameters()) + list(self.linear_layer.parameters()), lr=1e-5,
                                betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)          # :167-175
        n_steps = min(int(args.num_inference_steps / 5), len(pipeline.scheduler.alphas_cumprod))   # :198
        n_steps = max(n_steps, 1)
        history, step = [], 0
        for epoch in range(int(getattr(args, "num_epochs", 10))):                         # :122,202
            unet.train()
            total = torch.zeros((), device=self.device)
            batches = 0
            for image, _, label in train_loader:                                          # :209
                total += self.warmup_step(image, label, pipeline, unet, self.vae, self.linear_layer, opt, n_steps,
                                          seed=int(getattr(args, "seed", 0)), step_index=step)
                step += 1
                batches += 1
            mean = total / max(batches, 1)
            if dist.is_available() and dist.is_initialized():                             # _reduce_mean :56-64
                dist.all_reduce(mean)
                mean /= dist.get_world_size()
            history.append(float(mean))
            if log is not None:
                log.append((epoch + 1, history[-1]))
        return history

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
