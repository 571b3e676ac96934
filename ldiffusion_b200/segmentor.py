"""``Segmentor`` with the reference's interface (``segmentor.py:17-545``), its hot loops
running on the B200 kernels.

Drop-in surface kept: ``Segmentor(train_loader, val_loader, level, num_classes)``,
``ldiffusion_augment(inputs, pipeline, unet, vae)``,
``ldiffusion_augment_for_multimodal(rgb, dtm, pipeline, unet, vae, controlnet, batch_size, device)``,
``micro_dice(...)``,
``inference_cell_model(...)``, ``inference_tissue_model_nnUNetv2(...)``,
``load_ldiffusion(...)``, same ``ValueError``s.  The six copies of the sampling loop
in the reference are one method here (``_sample_and_decode``): scheduler
step = one fused launch, decode tail = one launch producing the uint8 image on the
device, no per-step device->host copy.

Out of scope and therefore injected: the SD pipeline / fine-tuned UNet loader
(diffusers), Cellpose + ResNet-152 instance features, nnU-Net.  ``pipeline_loader``
and ``model_factory`` default to the stand-ins of ``ldiffusion_b200.standin`` when
the real packages are not importable.
"""
import os

import numpy as np
import torch
from PIL import Image

from . import metrics, ops
from .head import cell_mask, tissue_mask

_IMAGENET_MEAN = (0.485, 0.456, 0.406)
_IMAGENET_STD = (0.229, 0.224, 0.225)


def _to_tensor_1024(image: Image.Image, device, normalize: bool):
    """transforms.Resize((1024,1024)) + ToTensor (+ Normalize(ImageNet)) (segmentor.py:505-510)."""
    img = image.resize((1024, 1024), Image.BILINEAR)
    x = torch.from_numpy(np.asarray(img, dtype=np.uint8).copy()).to(device).permute(2, 0, 1).float().div_(255.0)
    if normalize:
        mean = torch.tensor(_IMAGENET_MEAN, device=device).view(3, 1, 1)
        std = torch.tensor(_IMAGENET_STD, device=device).view(3, 1, 1)
        x = (x - mean) / std
    return x.unsqueeze(0).contiguous()


class Segmentor:
    _NO_LOADER = ("{what} is outside this package (the SD-v1.5 UNet / VAE and the segmentor networks are cuDNN library "
                  "calls, SURVEY 8): inject {arg}=..., or pass allow_standins=True to run the hot path on the "
                  "randomly initialised stand-ins of ldiffusion_b200.standin (results are then meaningless as masks)")

    def __init__(self, train_loader, val_loader, level, num_classes, pipeline_loader=None, model_factory=None,
                 allow_standins: bool = False):
        if not torch.cuda.is_available():
            raise RuntimeError("ldiffusion_b200.Segmentor needs a CUDA device (no CPU fallback)")
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.level = level
        self.num_classes = num_classes
        self.model = None
        self.train_loader, self.val_loader = train_loader, val_loader
        self.ldiffusion_proj = None
        self.pipeline_loader = pipeline_loader
        self.model_factory = model_factory
        self.allow_standins = bool(allow_standins)

    # -- out-of-scope pieces, injected ------------------------------------------
    def load_ldiffusion(self, ldiffusion_weight, diffusion_path):
        """segmentor.py:76-84 -> (pipeline, unet, vae).  The pipeline's scheduler is
        replaced by the product scheduler (same duck-type)."""
        if self.pipeline_loader is not None:
            pipeline, unet, vae = self.pipeline_loader(ldiffusion_weight, diffusion_path)
        elif not self.allow_standins:
            raise RuntimeError(self._NO_LOADER.format(what="loading the diffusion pipeline", arg="pipeline_loader"))
        else:
            from .standin import StandInPipeline
            pipeline = StandInPipeline(self.device)
            unet, vae = pipeline.unet, pipeline.vae
        from .scheduler import LaplacePLMSScheduler
        if not isinstance(pipeline.scheduler, LaplacePLMSScheduler):
            pipeline.scheduler = LaplacePLMSScheduler()
        return pipeline, unet, vae

    def initialize_model(self, level, num_classes, segmentor_weight=None):
        """segmentor.py:62-74, plus the weight loading the reference does right after it
        (``load_state_dict(torch.load(<segmentor_weight>/cellclassifier.pth))`` at segmentor.py:497; the nnU-Net
        predictor reads its own folder at :405/:465).  A ``model_factory`` that takes a third argument receives
        ``segmentor_weight`` and loads it itself; a two-argument factory gets the reference's cell-level load
        applied to what it returns."""
        if level not in ("tissue", "cell"):
            raise ValueError("Invalid level specified. Choose 'tissue' or 'cell'.")
        if self.model_factory is not None:
            import inspect
            try:
                takes_weight = len(inspect.signature(self.model_factory).parameters) >= 3
            except (TypeError, ValueError):
                takes_weight = False
            if takes_weight:
                return self.model_factory(level, num_classes, segmentor_weight)
            model = self.model_factory(level, num_classes)
            ckpt = os.path.join(str(segmentor_weight), "cellclassifier.pth") if segmentor_weight else None
            if level == "cell" and ckpt and os.path.exists(ckpt) and hasattr(model, "load_state_dict"):
                model.load_state_dict(torch.load(ckpt, weights_only=True))
            return model
        if not self.allow_standins:
            raise RuntimeError(self._NO_LOADER.format(what="building the segmentor network", arg="model_factory"))
        from .standin import StandInCellModel, StandInTissueModel
        return (StandInTissueModel if level == "tissue" else StandInCellModel)(num_classes, device=self.device)

    def _ensure_ldiffusion_proj(self, pipeline, unet, ldiffusion_weight=None):
        """segmentor.py:31-52 (the 768->cross_attention_dim text projection)."""
        hid = pipeline.text_encoder.config.hidden_size
        cad = unet.config.cross_attention_dim
        if self.ldiffusion_proj is None or self.ldiffusion_proj.in_features != hid \
                or self.ldiffusion_proj.out_features != cad:
            self.ldiffusion_proj = torch.nn.Linear(hid, cad).to(self.device, dtype=torch.float32)
        self.ldiffusion_proj.eval()
        return self.ldiffusion_proj

    @torch.no_grad()
    def _get_text_embeddings(self, prompt, batch_size, pipeline, unet):
        """segmentor.py:54-60."""
        proj = self._ensure_ldiffusion_proj(pipeline, unet)
        ids = pipeline.tokenizer([prompt] * batch_size)["input_ids"]
        ids = torch.tensor(ids, dtype=torch.long, device=self.device)
        return proj(pipeline.text_encoder(ids)["last_hidden_state"].to(torch.float32)).to(torch.float32)

    # -- the hot loop (segmentor.py:96-107, :436-448, :518-530; utils.py:189-205) ----
    @torch.no_grad()
    def _sample_and_decode(self, image, pipeline, unet, vae, text_embeddings, num_steps: int = 1,
                           model_input: bool = False):
        """image [B,3,H,W] -> uint8 RGB [B,H,W,3] on the device (what numpy_to_pil would hold);
        with ``model_input`` also the segmentor's normalised input tensor, fused into the same
        kernel (the reference goes decoded -> .cpu() -> PIL -> ToTensor -> Normalize -> device)."""
        sched = pipeline.scheduler
        latents = vae.encode(image).latent_dist.mean.to(dtype=torch.float32).contiguous()
        sched.set_timesteps(num_steps, device=self.device)
        rgb, mi = None, None
        n = len(sched.timesteps)
        for i, t in enumerate(sched.timesteps):
            latents = sched.scale_model_input(latents, t)
            output = unet(latents, t, text_embeddings)
            latents = sched.step(output[0].contiguous(), t, latents).prev_sample
            decoded = vae.decode((1 / 0.18215) * latents).sample      # decode_latents' first half: `1 / 0.18215 * latents`
            if model_input and i == n - 1:
                rgb, _, mi = ops.decode_tail_model_input(decoded.contiguous())
            else:
                rgb, _ = ops.decode_tail_gray(decoded.contiguous(), want_gray=False)
        return (rgb, mi) if model_input else rgb

    def ldiffusion_augment(self, inputs, pipeline, unet, vae):
        """segmentor.py:86-112: [B,3,H,W] -> float [B,3,1024,1024] on the device (the
        reference goes through PIL + Resize(1024) + ToTensor per image)."""
        self._ensure_ldiffusion_proj(pipeline, unet)
        text = self._get_text_embeddings("A pathological slide", 1, pipeline, unet)
        outs = []
        for index in range(len(inputs)):
            image = inputs[index].unsqueeze(0).to(self.device, dtype=torch.float32)
            rgb = self._sample_and_decode(image, pipeline, unet, vae, text)          # uint8 [1,H,W,3]
            x = rgb.permute(0, 3, 1, 2).float().div_(255.0)                          # ToTensor
            if x.shape[-2:] != (1024, 1024):                                         # Resize((1024,1024)), bilinear
                pil = Image.fromarray(rgb[0].cpu().numpy()).resize((1024, 1024), Image.BILINEAR)
                x = torch.from_numpy(np.asarray(pil).copy()).to(self.device).permute(2, 0, 1).float().div_(255.0)[None]
            outs.append(x)
        return torch.cat(outs, dim=0)

    @torch.no_grad()
    def ldiffusion_augment_for_multimodal(self, rgb, dtm, pipeline, unet, vae, controlnet, batch_size, device,
                                          *, noise=None, seed: int = 0):
        """segmentor.py:301-386: RGB + depth (DTM) -> list of fp32 ``[256,256,3]`` reconstructions.

        The Laplace(0,1) noise modulated by the resized depth map (``:339-345``) and its removal
        with the UNet's prediction (``:375,379``) are one launch each, for the whole batch, with the
        VAE's 0.18215 factors fused in; the depth map stays ``[B,1,h,w]`` and is broadcast over the
        latent channels inside the kernels (the reference materialises ``.repeat(1, C, 1, 1)`` and
        loops over the images with B = 1).  ControlNet / UNet / VAE are out of scope (injected).
        ``noise`` injects the Laplace(0,1) tensor (parity); otherwise Philox(``seed``)."""
        device = torch.device(device)
        self._ensure_ldiffusion_proj(pipeline, unet)
        rgb = ops.bilinear_lift(rgb.to(device, torch.float32).contiguous(), (256, 256))      # :322-323
        dtm = ops.bilinear_lift(dtm.to(device, torch.float32).contiguous(), (256, 256))      # :324-325
        B = dtm.shape[0]
        depth_condition = dtm.expand(B, 3, 256, 256)                                          # :335
        latents = vae.encode(rgb).latent_dist.sample().to(torch.float32).contiguous()        # :339 (x 0.18215 below)
        depth_resized = ops.bilinear_lift(dtm, tuple(latents.shape[-2:]))                    # :340, [B,1,32,32]
        latents_noisy = ops.laplace_qsample_map(latents, depth_resized, x_mul=0.18215, noise=noise,
                                                seed=seed)                                   # :339,344-345
        text = self._get_text_embeddings("A remote sense image", B, pipeline, unet)          # :348-351
        pipeline.scheduler.set_timesteps(1, device=device)                                   # :354
        noise_pred = None
        for timestep in pipeline.scheduler.timesteps:
            down, mid = controlnet(sample=latents_noisy, timestep=timestep, encoder_hidden_states=text,
                                   controlnet_cond=depth_condition, return_dict=False)       # :357-363
            noise_pred = unet(latents_noisy, timestep, encoder_hidden_states=text,
                              down_block_additional_residuals=down,
                              mid_block_additional_residual=mid).sample                      # :366-372
        z = ops.scaled_residual(latents_noisy, noise_pred.to(torch.float32).contiguous(), depth_resized,
                                out_div=0.18215)                                             # :375,379
        recon = vae.decode(z).sample                                                         # :379
        recon_np = recon.permute(0, 2, 3, 1).contiguous().cpu().numpy()                      # :381-383, one copy
        return [recon_np[i] for i in range(B)]

    def micro_dice(self, predicted_labels, true_labels, num_classes=7):
        """segmentor.py:114-142 (identical to utils.micro_dice)."""
        return metrics.micro_dice(predicted_labels, true_labels, num_classes)

    # -- inference entry points ---------------------------------------------------------
    def _load_rgb(self, image_path):
        image = Image.open(image_path).convert("RGB")
        if image.mode != "RGB":
            raise ValueError(f"Input image is not in RGB mode: {image.mode}")
        return image

    @torch.no_grad()
    def inference_cell_model(self, image_path, diffusion_path, ldiffusion_weight, segmentor_weight):
        """segmentor.py:490-545 -> (PIL decoded image, uint8 mask [H0,W0])."""
        if self.model is None:
            self.model = self.initialize_model("cell", self.num_classes, segmentor_weight)
        pipeline, unet, vae = self.load_ldiffusion(ldiffusion_weight, diffusion_path)
        image = self._load_rgb(image_path)
        width, height = image.size
        x = _to_tensor_1024(image, self.device, normalize=True)
        if x.dim() != 4:
            raise ValueError(f"Input image tensor has invalid dimensions: {x.dim()} (expected 4).")
        text = self._get_text_embeddings("A pathological slide", 1, pipeline, unet)
        rgb, mi = self._sample_and_decode(x, pipeline, unet, vae, text, model_input=True)   # uint8 [1,1024,1024,3]
        decoded_image = Image.fromarray(rgb[0].cpu().numpy())
        # model input: Resize/ToTensor/Normalize of the decoded image (HWC numpy in the reference);
        # Resize is the identity when the decode already is 1024x1024, so the fused tensor is exact
        if tuple(mi.shape[-2:]) != (1024, 1024):
            mi = _to_tensor_1024(decoded_image, self.device, normalize=True)
        model_input = mi[0].permute(1, 2, 0)
        inst_map, feats, ids = self.model.instances(model_input.cpu().numpy())
        clf = self.model.classifier
        mask = cell_mask(inst_map, feats.contiguous(), clf.weight.detach().contiguous(), clf.bias.detach(), ids,
                         lut_size=int(inst_map.max().item()) + 1)[0]
        ops.check_status(self.device)
        pred = Image.fromarray(mask.cpu().numpy()).resize((width, height), resample=Image.NEAREST)
        return decoded_image.resize((width, height), Image.BILINEAR), np.array(pred)

    @torch.no_grad()
    def inference_tissue_model_nnUNetv2(self, image_path, diffusion_path, ldiffusion_weight, segmentor_weight,
                                        output_path=None):
        """segmentor.py:388-488 front half + the TissueSegNet head path (conductor.py:127,135 +
        segmentor.py:536).  The nnU-Net predictor of the reference is out of scope; the mask comes
        from ``self.model.features`` + ``self.model.head`` through the fused head kernels."""
        if self.model is None:
            self.model = self.initialize_model("tissue", self.num_classes, segmentor_weight)
        pipeline, unet, vae = self.load_ldiffusion(ldiffusion_weight, diffusion_path)
        image = self._load_rgb(image_path)
        width, height = image.size
        if width == height:                                                          # segmentor.py:427
            x = _to_tensor_1024(image, self.device, normalize=True)
            text = self._get_text_embeddings("A pathological slide", 1, pipeline, unet)
            rgb, xin = self._sample_and_decode(x, pipeline, unet, vae, text, model_input=True)
            decoded_image = Image.fromarray(rgb[0].cpu().numpy())
            if tuple(xin.shape[-2:]) != (1024, 1024):
                xin = _to_tensor_1024(decoded_image, self.device, normalize=True)
        else:
            decoded_image = image
            xin = _to_tensor_1024(decoded_image, self.device, normalize=True)
        feat = self.model.features(xin).contiguous()
        head = self.model.head
        K = head.weight.shape[0]
        mask = tissue_mask(feat, head.weight.detach().reshape(K, -1).to(feat.dtype).contiguous(),
                           head.bias.detach(), xin.shape[2:])[0]
        pred = Image.fromarray(mask.cpu().numpy()).resize((width, height), resample=Image.NEAREST)
        return decoded_image.resize((width, height), Image.BILINEAR), np.array(pred)
