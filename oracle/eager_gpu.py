"""The reference's eager op chains of the hot path on CUDA tensors: the "PyTorch-eager on the same B200"
baseline SURVEY.md 2c / BASELINE.md name (``bench.py``'s ``eager_gpu_baseline`` leg), and the source of the
reference-side masks for the mask-mismatch report.

Test / bench infrastructure only (see ``oracle/__init__.py``).  Same stages, inputs and outputs as
``oracle/pipeline.py::run_chain``; everything stays on the device the inputs live on.  Where the reference
leaves the GPU (diffusers' ``decode_latents`` ends in ``.cpu().numpy()``, PIL does the gray conversion, the
metrics call ``.item()`` per class) this chain keeps the SAME arithmetic as device tensor ops — the most
favourable eager formulation, not a transfer benchmark:

* decode tail   ``(img / 2 + 0.5).clamp(0, 1)`` -> ``permute`` -> ``* 255`` -> ``round`` -> ``uint8``
                (torch.round is half-to-even like numpy's), PIL's 16.16 fixed-point luma as int32 ops
                [pixel_latent_vector.py:80-93]
* q_sample      ``torch.distributions.Laplace(0, b).sample`` + add                 [ldiffusion.py:233-237]
* scheduler     the PNDM/PLMS update as 0-dim-tensor arithmetic, op by op          [segmentor.py:100-104]
* features      ``F.interpolate(bilinear)`` -> weighted gray -> ``torch.cat``      [ldiffusion.py:224-226,240-251]
* tissue head   ``F.conv2d`` 1x1 -> ``F.interpolate`` -> ``softmax`` -> ``argmax`` [conductor.py:127,135; segmentor.py:536]
* cell head     ``F.linear`` -> ``softmax[:, 1:]`` -> ``topk`` -> LUT paint (the per-instance painting loop of
                conductor.py:224-231 is replaced by one indexing op: its literal form allocates a [1,K,H,W]
                tensor per instance and would measure the allocator)
* metrics       the confusion counts by ``torch.bincount`` (the reference's K^2 masked sums with ``.item()``
                per class are timed by the CPU leg; here the eager GPU gets the single-pass formulation)
"""
import torch
import torch.nn.functional as F

from .scheduler import PNDMOracle


def decode_tail(image):
    """[B,3,H,W] -> (uint8 [B,H,W,3], uint8 gray [B,H,W])."""
    x = (image / 2 + 0.5).clamp(0, 1)
    x = x.permute(0, 2, 3, 1).float()
    rgb = (x * 255).round().to(torch.uint8)
    r, g, b = (rgb[..., c].to(torch.int32) for c in range(3))
    gray = ((19595 * r + 38470 * g + 7471 * b + 0x8000) >> 16).to(torch.uint8)
    return rgb, gray


def confusion(mask, gt, K):
    """int64 [(K+1), K]: rows = gt class (row K = gt outside [0, K)), columns = predicted class."""
    g = torch.clamp(gt.to(torch.int64), max=K)
    return torch.bincount((g * K + mask.to(torch.int64)).view(-1), minlength=(K + 1) * K).view(K + 1, K)


@torch.no_grad()
def run_chain(inp, num_classes, head_w, head_b, cell_w, cell_b, feat_size=(64, 64), compute_dtype=torch.float32,
              ieee_fp32=False):
    """inp: HotPathInputs on a CUDA device.  ``compute_dtype``: fp32 is the reference's precision
    (``ldiffusion.py:67``); bf16 runs the same chain in the benchmark's storage type.  ``ieee_fp32``: switch TF32
    off for the convolution / linear (torch's CUDA default allows TF32 in cuDNN convolutions, which is what a user
    of the reference gets and what the timed baseline runs; the mask-mismatch reference wants true fp32 logits)."""
    if ieee_fp32:
        old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
        torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
        try:
            return run_chain(inp, num_classes, head_w, head_b, cell_w, cell_b, feat_size, compute_dtype, False)
        finally:
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    K, n = num_classes, len(inp.eps)
    dev = inp.latents.device
    cast = lambda t: t.to(compute_dtype)                                  # noqa: E731
    sch = PNDMOracle()
    sch.set_timesteps(n - 1)
    sch.alphas_cumprod = sch.alphas_cumprod.to(dev)                       # (the reference's table sits on the CPU and
    sch.final_alpha_cumprod = sch.final_alpha_cumprod.to(dev)             #  costs a sync per step: not charged here)
    x = cast(inp.latents)
    B, _, H, W = inp.decoded[0].shape
    out = {"noisy": [], "lat": []}
    grays, feat, rgb_u8, small = [], None, None, None
    for i, t in enumerate(sch.timesteps.tolist()):
        b_t = torch.sqrt(1 - sch.alphas_cumprod[t])
        noise = torch.distributions.Laplace(torch.zeros((), device=dev), b_t).sample(inp.latents.shape)
        out["noisy"].append(cast(inp.latents) + noise.to(compute_dtype))
        x = sch.step(cast(inp.eps[i]), t, x)
        out["lat"].append(x)
        rgb_u8, gray = decode_tail(inp.decoded[i])
        grays.append(gray)
        small = F.interpolate(cast(inp.decoded[i]), size=feat_size, mode="bilinear", align_corners=False)
        wts = torch.tensor([0.2989, 0.5870, 0.1140], device=dev, dtype=small.dtype).view(1, 3, 1, 1)
        gw = (small * wts).sum(dim=1, keepdim=True)
        feat = gw if feat is None else torch.cat([feat, gw], dim=1)
    gt = inp.gt
    out["pixel_planes"] = torch.stack(grays + [gt], dim=1)
    out["rgb"] = rgb_u8
    out["featcat"] = feat
    out["label_small"] = F.interpolate(gt.unsqueeze(1).to(torch.float32), size=feat_size, mode="bilinear",
                                       align_corners=False).to(torch.uint8)
    out["rgb_up"] = F.interpolate(small, size=(H, W), mode="bilinear", align_corners=False)
    # tissue head
    logits = F.conv2d(cast(inp.head_feat), cast(head_w)[:, :, None, None], cast(head_b))
    full = F.interpolate(logits, size=(H, W), mode="bilinear", align_corners=False)
    out["logits"] = logits
    out["mask_tissue"] = torch.argmax(torch.softmax(full, dim=1), dim=1).to(torch.uint8)
    out["full_logits"] = full
    # cell head
    n_inst = inp.inst_feats.shape[1]
    cl = F.linear(cast(inp.inst_feats), cast(cell_w), cast(cell_b))      # [B,N,K]
    probs = F.softmax(cl, dim=2)[:, :, 1:]
    cls = torch.topk(probs, k=1, dim=2)[1].squeeze(2) + 1                # [B,N] in 1..K-1
    lut = torch.zeros((B, n_inst + 1), dtype=torch.uint8, device=dev)
    lut[:, 1:] = cls.to(torch.uint8)
    out["cell_logits"] = cl
    out["mask_cell"] = torch.gather(lut, 1, inp.inst_map.view(B, -1).to(torch.int64)).view(B, H, W)
    out["confusion"] = torch.stack([confusion(out["mask_tissue"], gt, K), confusion(out["mask_cell"], gt, K)])
    return out
