"""Segmentation metrics from ONE confusion-matrix pass on the GPU.

Drop-in for ``utils.micro_dice`` (``utils.py:55-82`` == ``segmentor.py:114-142``),
``utils.mean_iou_and_per_class`` (``utils.py:84-104``), ``evaluate.pixel_accuracy``
(``evaluate.py:11-27``), ``evaluate.frequency_weighted_iou`` (``evaluate.py:29-45``)
and ``evaluate.evaluate`` (``evaluate.py:48-126``): same signatures, same return
types, same numbers.  The reference makes ~K^2+8K masked-sum passes with a host
sync each; here the label maps are read once by ``ldiff_confusion_hist`` into an
int64 ``[(K+1), K]`` matrix (row K collects gt values outside [0,K), which the
reference counts as false positives but never as a class), and every metric is
derived from that matrix on the host with the reference's own fp32 / python-float
arithmetic order.  The matrix is additive, so sharded evaluation is one integer
all-reduce (``ldiffusion_b200.dist``).
"""
import datetime
import glob
import os
from typing import Optional

import numpy as np
import torch

from . import ops


# ---------------------------------------------------------------------------
# device side
# ---------------------------------------------------------------------------

def _pred_labels_u8(pred: torch.Tensor) -> torch.Tensor:
    """The reference's metric functions take one-hot / logit maps [B,K,H,W] and
    argmax them; integer label maps are accepted directly."""
    if pred.is_floating_point():
        if pred.dim() < 3:
            raise ValueError("floating predictions must be [B,K,...]")
        return ops.argmax_channels(pred.contiguous())
    return ops.labels_to_u8(pred)


def confusion_matrix(pred: torch.Tensor, target: torch.Tensor, num_classes: int,
                     out: Optional[torch.Tensor] = None, gt_lut: Optional[torch.Tensor] = None) -> torch.Tensor:
    """int64 [(K+1),K] on the device; accumulates into ``out`` when given."""
    p = _pred_labels_u8(pred)
    g = ops.labels_to_u8(target)
    if p.numel() != g.numel():
        raise ValueError("prediction and target must have the same number of pixels")
    return ops.confusion_hist(p.reshape(-1), g.reshape(-1), num_classes, out=out, gt_lut=gt_lut)


def _to_host(C: torch.Tensor) -> np.ndarray:
    if C.is_cuda:
        ops.check_status(C.device)          # raises what F.one_hot would have raised
    return C.cpu().numpy()


# ---------------------------------------------------------------------------
# host epilogue: the reference's formulas evaluated on counts
# ---------------------------------------------------------------------------

def dice_from_confusion(C: np.ndarray):
    """utils.py:67-80.  TP/FP/FN are exact integers here; the reference obtains
    them as fp32 sums of 0/1 values, identical below 2**24 pixels per call."""
    K = C.shape[1]
    col = C.sum(0)
    dice = torch.zeros(K)
    for c in range(K):
        n_gt, n_pred, tp_i = int(C[c].sum()), int(col[c]), int(C[c, c])
        if n_gt == 0 and n_pred == 0:
            dice[c] = 1
            continue
        tp = torch.tensor(float(tp_i), dtype=torch.float32)
        fp = torch.tensor(float(n_pred - tp_i), dtype=torch.float32)
        fn = torch.tensor(float(n_gt - tp_i), dtype=torch.float32)
        dice[c] = 2 * tp / (2 * tp + fp + fn)
    return dice, torch.mean(dice)


def iou_from_confusion(C: np.ndarray):
    """utils.py:88-104."""
    K = C.shape[1]
    col = C.sum(0)
    iou_dict, ious = {}, []
    for c in range(K):
        inter = int(C[c, c])
        union = int(C[c].sum()) + int(col[c]) - inter
        if union == 0:
            iou_dict[c] = None
            continue
        iou_dict[c] = inter / union
        ious.append(iou_dict[c])
    return (sum(ious) / len(ious) if ious else 1.0), iou_dict


def pa_from_confusion(C: np.ndarray):
    """evaluate.py:15-27."""
    K = C.shape[1]
    acc = []
    for c in range(K):
        total = int(C[c].sum())
        acc.append(1.0 if total == 0 else int(C[c, c]) / total)
    return sum(acc) / len(acc), acc


def fwiou_from_confusion(C: np.ndarray, ignore_background: bool = False) -> float:
    """evaluate.py:32-44 (the 'other' gt row never enters ``hist``)."""
    K = C.shape[1]
    hist = torch.from_numpy(C[:K].astype(np.float32))
    freq = hist.sum(1) / hist.sum()
    diag = torch.diag(hist)
    iu = diag / (hist.sum(1) + hist.sum(0) - diag + 1e-10)
    if ignore_background:
        freq, iu = freq[1:], iu[1:]
    return (freq * iu).sum().item()


# ---------------------------------------------------------------------------
# drop-in signatures
# ---------------------------------------------------------------------------

def micro_dice(predicted_labels, true_labels, num_classes=7):
    C = _to_host(confusion_matrix(predicted_labels, true_labels, num_classes))
    dice, avg = dice_from_confusion(C)
    return dice.to(true_labels.device), avg.to(true_labels.device)


def mean_iou_and_per_class(pred, target, num_classes):
    return iou_from_confusion(_to_host(confusion_matrix(pred, target, num_classes)))


def pixel_accuracy(pred, target, num_classes):
    return pa_from_confusion(_to_host(confusion_matrix(pred, target, num_classes)))


def frequency_weighted_iou(pred, target, num_classes, ignore_background=False):
    return fwiou_from_confusion(_to_host(confusion_matrix(pred, target, num_classes)), ignore_background)


def image_metrics_from_confusion(C: np.ndarray):
    """The four foreground numbers evaluate.py:72-93 records for one image."""
    K = C.shape[1]
    dice, _ = dice_from_confusion(C)
    fg = dice[1:]
    _, iou = iou_from_confusion(C)
    vals = [iou[c] for c in range(1, K) if iou.get(c) is not None]
    _, pa = pa_from_confusion(C)
    return {
        "dice": torch.mean(fg).item(), "pc_dice": fg.numpy(),
        "iou": sum(vals) / len(vals) if vals else 1.0,
        "pc_iou": [iou[c] if iou.get(c) is not None else 1.0 for c in range(1, K)],
        "pa": np.mean(pa[1:]), "pc_pa": pa[1:],
        "fwiou": fwiou_from_confusion(C, ignore_background=True),
    }


def summarize_images(per_image_C):
    """evaluate.py:95-102: per-image metrics, then np.mean over images."""
    rows = [image_metrics_from_confusion(np.asarray(C)) for C in per_image_C]
    return {
        "mean_dice": np.mean([r["dice"] for r in rows]), "mean_iou": np.mean([r["iou"] for r in rows]),
        "mean_pa": np.mean([r["pa"] for r in rows]), "mean_fwiou": np.mean([r["fwiou"] for r in rows]),
        "per_class_dice": np.mean([r["pc_dice"] for r in rows], axis=0),
        "per_class_iou": np.mean([r["pc_iou"] for r in rows], axis=0),
        "per_class_pa": np.mean([r["pc_pa"] for r in rows], axis=0),
    }


# ---------------------------------------------------------------------------
# widening N4: nnU-Net's online validation counts and their collective
# ---------------------------------------------------------------------------

def tp_fp_fn_from_confusion(C: np.ndarray):
    """nnU-Net ``get_tp_fp_fn_tn`` on hard one-hot predictions, summed over batch and space
    (``nnUNetTrainer.py:954-986``; ``loss/dice.py:122-180``), from the confusion matrix.
    gt values outside [0,K) (the ignore label) are masked out, as nnU-Net's ``mask`` does.
    Returns float32 (tp, fp, fn) of length K-1 (background dropped, ``:982-984``)."""
    K = C.shape[1]
    inside = C[:K]
    tp = np.diag(inside).astype(np.float32)
    fp = (inside.sum(0) - np.diag(inside)).astype(np.float32)
    fn = (inside.sum(1) - np.diag(inside)).astype(np.float32)
    return tp[1:], fp[1:], fn[1:]


def online_tp_fp_fn(output: torch.Tensor, target: torch.Tensor, allreduce: bool = False):
    """Drop-in for the tail of ``nnUNetTrainer.validation_step`` (``:954-986``) plus, with
    ``allreduce``, the three ``all_gather_object`` calls of ``on_validation_epoch_end``
    (``:1004-1012``) as ONE int64 all-reduce.  output: logits [B,K,...]; target: labels
    [B,1,...] or [B,...].  Returns float32 numpy (tp_hard, fp_hard, fn_hard)."""
    K = output.shape[1]
    C = confusion_matrix(output, target.reshape(output.shape[0], *output.shape[2:]), K)
    if allreduce:
        from .dist import allreduce_confusion
        allreduce_confusion(C)
    return tp_fp_fn_from_confusion(_to_host(C))


def global_dice_from_counts(tp, fp, fn):
    """``nnUNetTrainer.py:1021-1022``: per-class pseudo Dice and its nanmean."""
    per_class = [i for i in [2 * i / (2 * i + j + k) for i, j, k in zip(tp, fp, fn)]]
    return per_class, np.nanmean(per_class)


def nnunet_label_metrics_from_confusion(C: np.ndarray, labels, ignore_label=None) -> dict:
    """nnU-Net's per-file metrics (``evaluation/evaluate_predictions.py:77-120``: ``compute_tp_fp_fn_tn`` +
    ``compute_metrics``) for plain integer labels, from the ``(K+1) x K`` confusion matrix (rows = reference
    label, row K = every reference value outside ``[0, K)``; columns = predicted label).

    ``ignore_label``: reference pixels with that value are left out of every count (``:97,79-82``); it is
    either a label below K or the only out-of-range value of the reference map (nnU-Net's convention: the
    ignore label is the highest one), in which case it is row K.  Returns ``{label: {"Dice", "IoU", "FP",
    "TP", "FN", "TN", "n_pred", "n_ref"}}`` with the reference's conventions (Dice = IoU = nan when
    tp + fp + fn == 0)."""
    C = np.asarray(C, dtype=np.int64)
    K = C.shape[1]
    used = np.ones(K + 1, dtype=bool)
    if ignore_label is not None:
        used[min(int(ignore_label), K)] = False
    Cu = C[used]
    total = int(Cu.sum())
    out = {}
    for r in labels:
        r = int(r)
        if not 0 <= r < K:
            raise ValueError(f"label {r} outside [0, {K})")
        tp = int(C[r, r]) if used[r] else 0
        fp = int(Cu[:, r].sum()) - tp
        fn = (int(C[r].sum()) - tp) if used[r] else 0
        tn = total - tp - fp - fn
        m = {}
        if tp + fp + fn == 0:
            m["Dice"], m["IoU"] = np.nan, np.nan
        else:
            m["Dice"], m["IoU"] = 2 * tp / (2 * tp + fp + fn), tp / (tp + fp + fn)
        m.update(FP=fp, TP=tp, FN=fn, TN=tn, n_pred=fp + tp, n_ref=fn + tp)
        out[r] = m
    return out


def nnunet_compute_metrics(seg_ref: torch.Tensor, seg_pred: torch.Tensor, labels, ignore_label=None) -> dict:
    """``compute_metrics`` of nnU-Net's evaluator on two label maps already on the device: one histogram
    launch instead of four boolean-mask reductions per label (``evaluate_predictions.py:97-119``)."""
    K = max(int(r) for r in labels) + 1
    C = confusion_matrix(seg_pred, seg_ref, K)
    return nnunet_label_metrics_from_confusion(_to_host(C), labels, ignore_label)


def per_image_confusion(preds: torch.Tensor, gts: torch.Tensor, num_classes: int) -> torch.Tensor:
    """uint8 [N,H,W] x2 -> int64 [N,(K+1),K] in one launch, no syncs."""
    return ops.confusion_hist_batched(preds, gts, int(num_classes))


def evaluate(image_dir, label_dir, num_classes, save_dir="./eval_results", device=None):
    """Drop-in for ``evaluate.evaluate`` (folder-vs-folder PNG scoring, same report file)."""
    from PIL import Image
    os.makedirs(save_dir, exist_ok=True)
    image_files = sorted(glob.glob(os.path.join(image_dir, "*.png")))
    label_files = sorted(glob.glob(os.path.join(label_dir, "*.png")))
    if len(image_files) != len(label_files):
        raise ValueError(f"The number of images: {len(image_files)}, The number of labels: "
                         f"{len(label_files)}, they must be equal.")
    device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    K = int(num_classes)
    mats = torch.zeros((len(image_files), K + 1, K), dtype=torch.int64, device=device)
    for i, (img_path, lbl_path) in enumerate(zip(image_files, label_files)):
        pred = np.array(Image.open(img_path))
        gt = np.array(Image.open(lbl_path))
        if pred.shape != gt.shape:
            raise ValueError(f"shape mismatch: {img_path} vs {lbl_path}")
        p = torch.from_numpy(np.ascontiguousarray(pred)).to(device, non_blocking=True)
        g = torch.from_numpy(np.ascontiguousarray(gt)).to(device, non_blocking=True)
        confusion_matrix(p, g, K, out=mats[i])
    s = summarize_images(_to_host(mats)) if len(image_files) else None
    timestamp = datetime.datetime.now().strftime("%Y%m%d_%H%M%S")
    save_path = os.path.join(save_dir, f"metrics_{timestamp}.txt")
    with open(save_path, "w") as f:
        f.write("=== Segmentation Evaluation Results ===\n")
        f.write(f"Image dir: {image_dir}\n")
        f.write(f"Label dir: {label_dir}\n")
        f.write(f"Classes: {num_classes}\n\n")
        f.write(f"The number of images: {len(image_files)}\n\n")
        if s is not None:
            f.write(f"Mean Dice:  {s['mean_dice']:.4f}\n")
            f.write(f"Mean IoU:   {s['mean_iou']:.4f}\n")
            f.write(f"Mean PA:    {s['mean_pa']:.4f}\n")
            f.write(f"Mean FWIoU: {s['mean_fwiou']:.4f}\n\n")
            f.write("Per-class metrics:\n")
            for c in range(1, K):
                f.write(f"Class {c}: Dice={s['per_class_dice'][c - 1]:.4f}, "
                        f"IoU={s['per_class_iou'][c - 1]:.4f}, PA={s['per_class_pa'][c - 1]:.4f}\n")
    print(f"Evaluation complete! Results saved to {save_path}")
    return s
