"""Oracle for stage a-6: confusion-matrix / Dice / IoU / PA / FWIoU reduction.

Test infrastructure only (see ``oracle/__init__.py``).

Two tiers:

* ``*_chain``: the reference's per-class masked-sum op chains, restated with
  torch-CPU ops (``utils.py:55-82``, ``utils.py:84-104``, ``evaluate.py:11-27``,
  ``evaluate.py:29-45``).  Pinned against the reference's own functions by
  ``tests/golden/make_golden.py``.
* ``confusion_matrix`` + ``*_from_confusion``: everything above is a function of
  ``C[i, j] = #(gt == i and pred == j)`` with one extra row ``i == K`` for gt
  values outside ``[0, K)`` (such pixels match no class in ``evaluate.py:35`` but
  still count as FP in ``utils.py:72`` and in the IoU union).  These are the
  bit-exact targets of the CUDA ``confusion_hist`` kernel and its host epilogue.
"""
import numpy as np
import torch


# --------------------------------------------------------------------------
# chain tier
# --------------------------------------------------------------------------

def micro_dice_chain(predicted_labels, true_labels, num_classes=7):
    """utils.py:55-82 (identical copy at segmentor.py:114-142).

    predicted_labels: float [B, K, H, W] (argmax over dim 1 is taken here),
    true_labels: integer tensor with B*H*W elements.
    Returns (dice fp32 [K], mean fp32 []).
    """
    pred = torch.argmax(predicted_labels, dim=1).reshape(-1)
    true = true_labels.reshape(-1)
    dice = torch.zeros(num_classes, device=true.device)
    for c in range(num_classes):
        t = (true == c).float()
        p = (pred == c).float()
        if torch.sum(t) == 0 and torch.sum(p) == 0:
            dice[c] = 1
            continue
        tp = torch.sum(t * p)
        fp = torch.sum((1 - t) * p)
        fn = torch.sum(t * (1 - p))
        if tp + fp + fn == 0:
            dice[c] = 0
        else:
            dice[c] = 2 * tp / (2 * tp + fp + fn)
    return dice, torch.mean(dice)


def mean_iou_and_per_class_chain(pred, target, num_classes):
    """utils.py:84-104.  Returns (float, {class: float | None})."""
    labels = torch.argmax(pred, dim=1)
    per_class = {}
    kept = []
    for c in range(num_classes):
        a = labels == c
        b = target == c
        inter = (a & b).sum().item()
        union = (a | b).sum().item()
        if union == 0:
            per_class[c] = None
        else:
            per_class[c] = inter / union
            kept.append(per_class[c])
    return (sum(kept) / len(kept) if kept else 1.0), per_class


def pixel_accuracy_chain(pred, target, num_classes):
    """evaluate.py:11-27.  Returns (float, list[float])."""
    labels = torch.argmax(pred, dim=1)
    acc = []
    for c in range(num_classes):
        b = target == c
        total = b.sum().item()
        if total == 0:
            acc.append(1.0)
        else:
            acc.append(((labels == c) & b).sum().item() / total)
    return sum(acc) / len(acc), acc


def frequency_weighted_iou_chain(pred, target, num_classes, ignore_background=False):
    """evaluate.py:29-45.  K*K masked sums into a float hist, then fp32 algebra."""
    labels = torch.argmax(pred, dim=1)
    hist = torch.zeros((num_classes, num_classes), dtype=torch.float)
    for i in range(num_classes):
        gi = target == i
        for j in range(num_classes):
            hist[i, j] = (gi & (labels == j)).sum().item()
    return _fwiou_from_hist(hist, ignore_background)


def _fwiou_from_hist(hist, ignore_background):
    # evaluate.py:37-44
    freq = hist.sum(1) / hist.sum()
    diag = torch.diag(hist)
    iu = diag / (hist.sum(1) + hist.sum(0) - diag + 1e-10)
    if ignore_background:
        freq = freq[1:]
        iu = iu[1:]
    return (freq * iu).sum().item()


# --------------------------------------------------------------------------
# confusion-matrix tier
# --------------------------------------------------------------------------

def confusion_matrix(pred, gt, num_classes):
    """int64 [(K+1), K]: rows = gt class (row K = gt outside [0,K)), cols = pred.

    pred values outside [0, K) raise, as ``F.one_hot`` does at evaluate.py:70.
    """
    K = int(num_classes)
    p = np.asarray(pred).reshape(-1).astype(np.int64)
    g = np.asarray(gt).reshape(-1).astype(np.int64)
    if p.shape != g.shape:
        raise ValueError("pred and gt must have the same number of pixels")
    if p.size and (p.min() < 0 or p.max() >= K):
        raise RuntimeError("Class values must be smaller than num_classes.")
    row = np.where((g >= 0) & (g < K), g, K)
    flat = np.bincount(row * K + p, minlength=(K + 1) * K)
    return flat.reshape(K + 1, K).astype(np.int64)


def micro_dice_from_confusion(C):
    """utils.py:67-80 evaluated from counts.  fp32 arithmetic in the chain's
    order: ``2*TP / ((2*TP + FP) + FN)``.  Counts below 2**24 are exact in the
    chain's fp32 sums, so the two tiers agree bit for bit there."""
    C = np.asarray(C, dtype=np.int64)
    K = C.shape[1]
    dice = torch.zeros(K)
    col = C.sum(0)
    for c in range(K):
        n_gt = int(C[c].sum())
        n_pred = int(col[c])
        if n_gt == 0 and n_pred == 0:
            dice[c] = 1
            continue
        tp = torch.tensor(float(C[c, c]), dtype=torch.float32)
        fp = torch.tensor(float(n_pred - C[c, c]), dtype=torch.float32)
        fn = torch.tensor(float(n_gt - C[c, c]), dtype=torch.float32)
        dice[c] = 2 * tp / (2 * tp + fp + fn)
    return dice, torch.mean(dice)


def mean_iou_from_confusion(C):
    """utils.py:88-104 from counts (python-int division, as ``.item()`` gives)."""
    C = np.asarray(C, dtype=np.int64)
    K = C.shape[1]
    col = C.sum(0)
    per_class = {}
    kept = []
    for c in range(K):
        inter = int(C[c, c])
        union = int(C[c].sum()) + int(col[c]) - inter
        if union == 0:
            per_class[c] = None
        else:
            per_class[c] = inter / union
            kept.append(per_class[c])
    return (sum(kept) / len(kept) if kept else 1.0), per_class


def pixel_accuracy_from_confusion(C):
    """evaluate.py:15-27 from counts."""
    C = np.asarray(C, dtype=np.int64)
    K = C.shape[1]
    acc = []
    for c in range(K):
        total = int(C[c].sum())
        acc.append(1.0 if total == 0 else int(C[c, c]) / total)
    return sum(acc) / len(acc), acc


def fwiou_from_confusion(C, ignore_background=False):
    """evaluate.py:32-44: the K*K block of C *is* ``hist`` (the "other" row is
    invisible to it)."""
    C = np.asarray(C, dtype=np.int64)
    K = C.shape[1]
    hist = torch.from_numpy(C[:K].astype(np.float32))
    return _fwiou_from_hist(hist, ignore_background)


def evaluate_images_chain(preds, gts, num_classes):
    """evaluate.py:60-102 without the file system: per-image metrics with the
    background dropped, then ``np.mean`` over images.  preds/gts: lists of 2-D
    integer arrays.  Returns a dict of the numbers evaluate() writes."""
    K = num_classes
    all_dice, all_iou, all_pa, all_fw = [], [], [], []
    pc_dice, pc_iou, pc_pa = [], [], []
    for pred, gt in zip(preds, gts):
        if np.shape(pred) != np.shape(gt):
            raise ValueError("shape mismatch")
        p = torch.from_numpy(np.asarray(pred)).long().unsqueeze(0)
        g = torch.from_numpy(np.asarray(gt)).long()
        onehot = torch.nn.functional.one_hot(p, num_classes=K).permute(0, 3, 1, 2).float()
        d, _ = micro_dice_chain(onehot, g, K)
        all_dice.append(torch.mean(d[1:]).item())
        pc_dice.append(d[1:].numpy())
        _, iou = mean_iou_and_per_class_chain(onehot, g, K)
        vals = [iou[c] for c in range(1, K) if iou.get(c) is not None]
        all_iou.append(sum(vals) / len(vals) if vals else 1.0)
        pc_iou.append([iou[c] if iou.get(c) is not None else 1.0 for c in range(1, K)])
        _, pa = pixel_accuracy_chain(onehot, g, K)
        all_pa.append(np.mean(pa[1:]))
        pc_pa.append(pa[1:])
        all_fw.append(frequency_weighted_iou_chain(onehot, g, K, ignore_background=True))
    return {
        "mean_dice": np.mean(all_dice), "mean_iou": np.mean(all_iou),
        "mean_pa": np.mean(all_pa), "mean_fwiou": np.mean(all_fw),
        "per_class_dice": np.mean(pc_dice, axis=0),
        "per_class_iou": np.mean(pc_iou, axis=0),
        "per_class_pa": np.mean(pc_pa, axis=0),
    }


# --------------------------------------------------------------------------
# widening N4: nnU-Net online validation counts (test infrastructure only)
# --------------------------------------------------------------------------

def nnunet_tp_fp_fn_chain(output, target, ignore_label=None):
    """nnUNetTrainer.py:954-986 restated: argmax -> one-hot prediction -> masked products with
    the one-hot target -> sums over batch and space; background dropped.  Pinned against the
    vendored ``get_tp_fp_fn_tn`` by tests/golden/nnunet_counts.npz."""
    K = output.shape[1]
    seg = output.argmax(1)[:, None]
    pred = torch.zeros(output.shape, dtype=torch.float32)
    pred.scatter_(1, seg, 1)
    tgt = target.clone()
    mask = None
    if ignore_label is not None:
        mask = (tgt != ignore_label).float()
        tgt[tgt == ignore_label] = 0
    y = torch.zeros(output.shape, dtype=torch.float32)
    y.scatter_(1, tgt.long(), 1)
    tp, fp, fn = pred * y, pred * (1 - y), (1 - pred) * y
    if mask is not None:
        tp, fp, fn = tp * mask, fp * mask, fn * mask
    axes = [0] + list(range(2, output.ndim))
    return tuple(t.sum(dim=axes).numpy()[1:] for t in (tp, fp, fn))
