"""The whole hot path on the CPU, stage for stage, with the reference's eager op
chains (test infrastructure + the timed CPU baseline of bench.py).

Mirrors ``ldiffusion_b200.pipeline.HotPath.run`` (same inputs, same outputs) so
that (i) tests can compare every output of one GPU pass against it and (ii)
``bench.py`` can time "the reference's PyTorch CPU path" on the box's host cores
for the same workload.  Reference lines are cited per stage in
``ldiffusion_b200/pipeline.py`` and in the sibling oracle modules.
"""
import numpy as np
import torch

from . import bilinear as obil
from . import decode_tail as odt
from . import head as ohead
from . import laplace as olap
from . import metrics as omet
from .scheduler import PNDMOracle


def run_chain(inp, num_classes, head_w, head_b, cell_w, cell_b, feat_size=(64, 64), noise=None,
              metrics="chain", paint="lut", storage=None, spec_lifts=False):
    """inp: an object with the fields of HotPathInputs holding CPU tensors (any
    float dtype; computed in fp32 as the reference does).  ``noise``: optional
    list of injected Laplace noise tensors (else torch's own sampler is used, as
    at ldiffusion.py:235-236).  Returns a dict of outputs.

    ``storage`` (e.g. torch.bfloat16) emulates a build that STORES every floating tensor it hands from one
    stage to the next in that type while computing in fp32: each stage's result is rounded once to
    ``storage`` (and widened again for the next stage), exactly what the CUDA pass does with bf16 buffers —
    the bf16 pass is then compared bit for bit instead of within a tolerance.  ``spec_lifts``: take the
    bilinear stages from the pinned-rounding spec tier (``oracle.bilinear.*_spec``) instead of ATen-CPU, whose
    down-sampling bits depend on its loop specialisation (<= 2 ulp apart)."""
    K = num_classes
    n = len(inp.eps)
    f32 = lambda t: t.to(torch.float32)                                   # noqa: E731
    st = (lambda t: t.to(storage).to(torch.float32)) if storage is not None else (lambda t: t)   # noqa: E731
    sch = PNDMOracle()
    sch.set_timesteps(n - 1)
    x = f32(inp.latents)
    B, _, H, W = inp.decoded[0].shape
    out = {"noisy": [], "lat": []}
    grays, feat = [], None
    rgb_u8 = None
    for i, t in enumerate(sch.timesteps):
        # a-1 (ldiffusion.py:233-237)
        if noise is not None:
            out["noisy"].append(st(olap.qsample_injected(f32(inp.latents), noise[i])))
        else:
            out["noisy"].append(st(olap.qsample_chain(f32(inp.latents), t)[0]))
        # a-2 (segmentor.py:102-104)
        x = st(sch.step(f32(inp.eps[i]), t, sch.scale_model_input(x, t)))
        out["lat"].append(x)
        # a-3 (pixel_latent_vector.py:80-86)
        rgb_u8 = odt.decode_tail_chain(inp.decoded[i])
        grays.append(odt.gray_chain(rgb_u8))
        # a-4 (ldiffusion.py:240-247)
        if spec_lifts:
            small = torch.from_numpy(obil.lift_spec(f32(inp.decoded[i]).numpy(), feat_size))
            gw = st(torch.from_numpy(obil.gray_weighted_spec(small.numpy())))
        else:
            small = obil.lift_chain(f32(inp.decoded[i]), feat_size)
            gw = st(obil.gray_weighted_chain(small))
        small = st(small)
        feat = gw if feat is None else torch.cat([feat, gw], dim=1)
    gt = inp.gt
    out["pixel_planes"] = np.stack(grays + [gt.numpy()], axis=1)          # [B,n+1,H,W]
    out["rgb"] = rgb_u8
    out["featcat"] = feat
    out["label_small"] = obil.label_down_chain(gt.unsqueeze(1), feat_size)  # ldiffusion.py:224-226
    out["rgb_up"] = st(torch.from_numpy(obil.lift_spec(small.numpy(), (H, W))) if spec_lifts
                       else obil.lift_chain(small, (H, W)))                 # ldiffusion.py:251
    # a-5 tissue (conductor.py:127,135 + segmentor.py:536)
    mask_t, logits = ohead.head_argmax_chain(f32(inp.head_feat), f32(head_w), f32(head_b), (H, W))
    out["logits"] = logits
    out["mask_tissue"] = mask_t.to(torch.uint8)
    # a-5 cell (conductor.py:218-231 + segmentor.py:536)
    n_inst = inp.inst_feats.shape[1]
    ids = np.arange(1, n_inst + 1)
    masks = []
    for b in range(B):
        cls, _ = ohead.cell_classify_chain(f32(inp.inst_feats[b]), f32(cell_w), f32(cell_b))
        if paint == "loop":         # the literal per-instance loop (allocates [1,K,H,W] per instance)
            masks.append(ohead.cell_paint_chain(inp.inst_map[b].numpy(), ids, cls.numpy(), K).to(torch.uint8))
        else:
            lut = ohead.cell_lut_spec(ids, cls.numpy(), n_inst + 1)
            masks.append(torch.from_numpy(ohead.cell_paint_spec(inp.inst_map[b].numpy(), lut)))
    out["mask_cell"] = torch.stack(masks)
    # a-6 (utils.py:55-104, evaluate.py:11-45) per image, as evaluate.py:60-93 does
    conf = []
    for m in (out["mask_tissue"], out["mask_cell"]):
        if metrics == "chain":
            rows = []
            for b in range(B):
                onehot = torch.nn.functional.one_hot(m[b].long().unsqueeze(0), K).permute(0, 3, 1, 2).float()
                g = gt[b].long()
                rows.append((omet.micro_dice_chain(onehot, g, K), omet.mean_iou_and_per_class_chain(onehot, g, K),
                             omet.pixel_accuracy_chain(onehot, g, K),
                             omet.frequency_weighted_iou_chain(onehot, g, K, True)))
            out.setdefault("metric_rows", []).append(rows)
        conf.append(omet.confusion_matrix(m.numpy(), gt.numpy(), K))
    out["confusion"] = np.stack(conf)
    return out
