"""Generate the golden fixtures under tests/golden/ (run in the build container).

    python tests/golden/make_golden.py

* metrics_*.npz  — inputs + outputs of the REFERENCE's own functions
  (utils.micro_dice, utils.mean_iou_and_per_class, evaluate.pixel_accuracy,
  evaluate.frequency_weighted_iou), imported from /root/reference through
  oracle/_refshim.py.  These pin the oracle's metric tier.
* chain_*.npz    — outputs of the third-party code the reference calls on this
  path, run here: torch.distributions.Laplace's transform, F.interpolate,
  argmax(softmax(.)), numpy round + PIL convert("L").  These pin the spec tier.

The reference cannot travel to the GPU box; the fixtures (a few hundred KB) do.
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import _refshim  # noqa: E402


def blob_labels(shape, K, seed, other_frac=0.001):
    rng = np.random.default_rng(seed)
    B, H, W = shape
    out = np.zeros(shape, np.uint8)
    for b in range(B):
        for _ in range(max(1, H * W // 600)):
            s = int(rng.integers(4, 24)); y = int(rng.integers(0, max(1, H - s))); x = int(rng.integers(0, max(1, W - s)))
            out[b, y:y + s, x:x + s] = rng.integers(1, K)
    out[rng.random(shape) < other_frac] = 255
    return out


def metrics_golden():
    ref_utils, ref_eval = _refshim.load()
    cases = [("k11", 11, (1, 96, 128), 0.002), ("k7", 7, (2, 64, 64), 0.0), ("k6_absent", 6, (1, 48, 80), 0.01),
             ("k11_allbg", 11, (1, 32, 32), 0.0)]
    for name, K, shape, other in cases:
        gt = blob_labels(shape, K, 100 + K, other)
        pred = blob_labels(shape, K, 200 + K, 0.0)
        if name == "k6_absent":
            pred[pred == 3] = 0; gt[gt == 3] = 0; gt[gt == 5] = 0      # class absent in both / in gt only
        if name == "k11_allbg":
            pred[:] = 0; gt[:] = 0
        p = torch.from_numpy(pred).long()
        g = torch.from_numpy(gt).long()
        onehot = F.one_hot(p, K).permute(0, 3, 1, 2).float()
        dice, avg = ref_utils.micro_dice(onehot, g, num_classes=K)
        miou, iou = ref_utils.mean_iou_and_per_class(onehot, g, K)
        mpa, pa = ref_eval.pixel_accuracy(onehot, g, K)
        fw0 = ref_eval.frequency_weighted_iou(onehot, g, K, ignore_background=False)
        fw1 = ref_eval.frequency_weighted_iou(onehot, g, K, ignore_background=True)
        np.savez_compressed(
            os.path.join(HERE, f"metrics_{name}.npz"), pred=pred, gt=gt, K=K,
            dice=dice.numpy(), avg_dice=avg.numpy(), miou=np.float64(miou),
            iou=np.array([np.nan if iou[c] is None else iou[c] for c in range(K)], np.float64),
            mpa=np.float64(mpa), pa=np.array(pa, np.float64), fwiou=np.float64(fw0), fwiou_fg=np.float64(fw1))
        print("metrics", name, float(avg), miou)


def chain_golden():
    g = torch.Generator().manual_seed(77)
    # Laplace transform of torch.distributions.Laplace (a-1)
    u = torch.empty(4096).uniform_(torch.finfo(torch.float32).eps - 1, 1, generator=g)
    b = torch.tensor(0.75968331)
    lap = torch.distributions.Laplace(0, b)
    noise = lap.loc - lap.scale * u.sign() * torch.log1p(-u.abs())
    # bilinear (a-4) and head decision (a-5)
    x_up = torch.randn(1, 3, 16, 16, generator=g)
    up = F.interpolate(x_up, size=(128, 128), mode="bilinear", align_corners=False)
    x_odd = torch.randn(1, 2, 13, 9, generator=g)
    odd = F.interpolate(x_odd, size=(31, 40), mode="bilinear", align_corners=False)
    x_dn = torch.randn(1, 3, 128, 128, generator=g)
    dn = F.interpolate(x_dn, size=(8, 8), mode="bilinear", align_corners=False)
    gray = (dn * torch.tensor([0.2989, 0.5870, 0.1140]).view(1, 3, 1, 1)).sum(dim=1, keepdim=True)
    lab = torch.randint(0, 256, (1, 1, 128, 128), generator=g, dtype=torch.uint8)
    lab_dn = F.interpolate(lab.float(), size=(8, 8), mode="bilinear", align_corners=False).to(torch.uint8)
    logits = torch.randn(1, 11, 8, 8, generator=g)
    mask = torch.argmax(torch.softmax(F.interpolate(logits, size=(256, 256), mode="bilinear",
                                                    align_corners=False), dim=1), dim=1)
    # decode tail + PIL gray (a-3)
    dec = torch.empty(1, 3, 64, 64).uniform_(-1.2, 1.2, generator=g)
    k = torch.arange(256)
    dec.view(-1)[:256] = (k.float() + 0.5) / 255.0 * 2 - 1          # exact .5 boundaries
    img = (dec / 2 + 0.5).clamp(0, 1).permute(0, 2, 3, 1).float().numpy()
    rgb = (img * 255).round().astype("uint8")
    pil_gray = np.array(Image.fromarray(rgb[0]).convert("L"))
    decb = dec.bfloat16()
    imgb = (decb / 2 + 0.5).clamp(0, 1).permute(0, 2, 3, 1).float().numpy()
    rgbb = (imgb * 255).round().astype("uint8")
    np.savez_compressed(
        os.path.join(HERE, "chain_ops.npz"), u=u.numpy(), b=b.numpy(), noise=noise.numpy(),
        x_up=x_up.numpy(), up=up.numpy(), x_odd=x_odd.numpy(), odd=odd.numpy(), x_dn=x_dn.numpy(), dn=dn.numpy(),
        gray=gray.numpy(), lab=lab.numpy(), lab_dn=lab_dn.numpy(), logits=logits.numpy(),
        mask=mask.numpy().astype(np.uint8), dec=dec.numpy(), rgb=rgb, pil_gray=pil_gray, rgb_bf16=rgbb)
    print("chain ops saved")


def nnunet_golden():
    """N4: the reference's vendored nnU-Net online validation counts (loss/dice.py:122-180 as
    called at nnUNetTrainer.py:954-986), with and without an ignore label."""
    sys.path.insert(0, os.path.join(_refshim.REF_ROOT, "model"))
    from nnunetv2.training.loss.dice import get_tp_fp_fn_tn
    g = torch.Generator().manual_seed(5)
    K = 5
    output = torch.randn(3, K, 40, 48, generator=g)
    target = torch.randint(0, K + 1, (3, 1, 40, 48), generator=g)          # value K plays the ignore label
    out = {}
    for name, ignore in (("plain", None), ("ignore", K)):
        tgt = target.clone()
        if ignore is None:
            tgt[tgt == K] = 0
        seg = output.argmax(1)[:, None]
        onehot = torch.zeros(output.shape, dtype=torch.float32)
        onehot.scatter_(1, seg, 1)
        mask = None
        if ignore is not None:
            mask = (tgt != ignore).float()
            tgt[tgt == ignore] = 0
        axes = [0] + list(range(2, output.ndim))
        tp, fp, fn, _ = get_tp_fp_fn_tn(onehot, tgt, axes=axes, mask=mask)
        out[name] = np.stack([tp.numpy()[1:], fp.numpy()[1:], fn.numpy()[1:]])
    np.savez_compressed(os.path.join(HERE, "nnunet_counts.npz"), output=output.numpy(), target=target.numpy(),
                        plain=out["plain"], ignore=out["ignore"], K=K)
    print("nnunet counts saved")


def nnunet_eval_golden():
    """Second pin of the confusion tier: nnU-Net's evaluator (evaluation/evaluate_predictions.py:67-119)."""
    path = os.path.join(_refshim.REF_ROOT, "model", "nnunetv2", "evaluation", "evaluate_predictions.py")
    to_mask, counts = _refshim.load_functions(path, ["region_or_label_to_mask", "compute_tp_fp_fn_tn"])
    rng = np.random.default_rng(77)
    K = 6
    ref = rng.integers(0, K + 1, (3, 40, 52)).astype(np.uint8)        # value K = the ignore label
    pred = rng.integers(0, K, (3, 40, 52)).astype(np.uint8)
    pred[ref == 2] = 2                                                # a perfectly segmented class
    ref[ref == 4] = 0; pred[pred == 4] = 0                            # a class absent from both
    out = {}
    for tag, ignore in (("plain", None), ("ignore", K), ("ignore_inner", 3)):
        ignore_mask = (ref == ignore) if ignore is not None else None
        out[tag] = np.array([counts(to_mask(ref, r), to_mask(pred, r), ignore_mask) for r in range(1, K)], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "nnunet_eval.npz"), ref=ref, pred=pred, K=K, **out)
    print("nnunet eval saved", out["plain"][0], out["ignore"][0])


def sliding_golden():
    """N2: the vendored nnU-Net sliding-window helpers (sliding_window_prediction.py:10-56); the
    module's unused acvl_utils import is stubbed."""
    import types
    sys.path.insert(0, os.path.join(_refshim.REF_ROOT, "model"))
    for name in ("acvl_utils", "acvl_utils.cropping_and_padding", "acvl_utils.cropping_and_padding.padding"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["acvl_utils.cropping_and_padding.padding"].pad_nd_image = None
    from nnunetv2.inference.sliding_window_prediction import compute_gaussian, compute_steps_for_sliding_window
    out = {}
    for ts in ((64, 64), (48, 80), (512, 512)):
        g = compute_gaussian(ts, sigma_scale=1. / 8, value_scaling_factor=10, dtype=torch.float16,
                             device=torch.device("cpu"))
        out[f"gauss_{ts[0]}x{ts[1]}"] = g.numpy() if ts[0] <= 80 else g.numpy()[::8, ::8].copy()
    cases = [((110, 110), (64, 64), 0.5), ((1024, 1024), (512, 512), 0.5), ((300, 260), (128, 96), 0.5),
             ((64, 64), (64, 64), 0.5), ((1000, 700), (256, 256), 0.25)]
    steps = [compute_steps_for_sliding_window(a, b, c) for a, b, c in cases]
    np.savez_compressed(os.path.join(HERE, "nnunet_sliding.npz"), steps=np.array(repr(steps)),
                        cases=np.array(repr(cases)), **out)
    print("nnunet sliding saved", steps[0], steps[2])


if __name__ == "__main__":
    if not _refshim.available():
        raise SystemExit("reference checkout not found; golden metric vectors need it")
    metrics_golden()
    chain_golden()
    nnunet_golden()
    nnunet_eval_golden()
    sliding_golden()
