"""torch-facing operators of the hot path.

Every operator is a ``torch.library`` custom op (namespace ``ldiff``) whose
implementation is one call into the C ABI of ``libldiff_sm100.so`` on the
caller's current CUDA stream.  Tensors must live on a CUDA device: there is no
CPU implementation and none is dispatched to.
"""
from typing import Optional, Sequence

import torch
from torch import Tensor

from . import _cabi
from ._cabi import BF16, F32, U8, LdiffError, check

_DT = {torch.float32: F32, torch.bfloat16: BF16, torch.uint8: U8}


def _dt(t: Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise TypeError(f"ldiff: unsupported dtype {t.dtype}") from None


def _cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise LdiffError("ldiff operators run on CUDA tensors only (no CPU fallback)")


def _ptr(t: Optional[Tensor]):
    return None if t is None else t.data_ptr()


def _stream(t: Tensor):
    return torch.cuda.current_stream(t.device).cuda_stream


def _dense(t: Tensor, what: str):
    if not t.is_contiguous():
        raise ValueError(f"ldiff: {what} must be contiguous")


_status_words = {}


def status_word(device) -> Tensor:
    """Per-device int32 word the kernels OR data-error bits into."""
    device = torch.device(device)
    key = device.index if device.index is not None else torch.cuda.current_device()
    w = _status_words.get(key)
    if w is None:
        w = torch.zeros(1, dtype=torch.int32, device=torch.device("cuda", key))
        _status_words[key] = w
    return w


def check_status(device):
    """Synchronising read of the status word; raises like the reference does.  Only the bit that is reported is
    cleared: a second error recorded at the same time surfaces on the next call instead of being dropped."""
    w = status_word(device)
    bits = int(w.item())
    if not bits:
        return

    def report(bit, exc):
        w.bitwise_and_(~bit)
        raise exc

    if bits & _cabi.STATUS_PRED_RANGE:
        # F.one_hot at evaluate.py:70 raises this for a predicted label >= K
        report(_cabi.STATUS_PRED_RANGE, RuntimeError("Class values must be smaller than num_classes."))
    if bits & _cabi.STATUS_INST_RANGE:
        report(_cabi.STATUS_INST_RANGE, RuntimeError("ldiff: instance id outside the class LUT"))
    if bits & _cabi.STATUS_SW_INF:
        # predict_from_raw_data.py:581-585
        report(_cabi.STATUS_SW_INF, RuntimeError(
            "Encountered inf in predicted array. Aborting... If this problem persists, "
            "reduce value_scaling_factor in compute_gaussian or increase the dtype of "
            "predicted_logits to fp32"))
    if bits & _cabi.STATUS_LABEL_RANGE:
        report(_cabi.STATUS_LABEL_RANGE, RuntimeError("ldiff: label value >= 32 in the contrastive sampler"))
    if bits & _cabi.STATUS_XCHG_TIMEOUT:
        report(_cabi.STATUS_XCHG_TIMEOUT, RuntimeError(
            "ldiff: a rank did not deliver its confusion matrix in time (peer exchange): the sums read since "
            "are partial"))
    w.zero_()                                          # unknown bits


# ---------------------------------------------------------------------------
# raw out-variant entry points: one ctypes call each.  They are registered as
# torch.library custom ops (torch.ops.ldiff.*) for callers that want a
# dispatcher-visible op; the functional wrappers below call them directly,
# because the dispatcher round trip costs ~25 us per call on the host and the
# latent-sized kernels run for 2-3 us.
# ---------------------------------------------------------------------------

def _laplace_qsample(x: Tensor, out: Tensor, noise: Optional[Tensor], u: Optional[Tensor],
                     noise_out: Optional[Tensor], b: float, seed: int, offset: int) -> None:
    _cuda(x, out, noise, u, noise_out)
    if x.numel() == 0:
        return
    check(_cabi.lib().ldiff_laplace_qsample(_ptr(x), _ptr(out), _ptr(noise), _ptr(u), _ptr(noise_out),
                                            b, seed, offset, x.numel(), _dt(x), _stream(x)))


def _map_dims(x: Tensor, scale: Tensor):
    """(plane, channels, scale_channels) of a [B,C,...] tensor and its [B,C,...] / [B,1,...] scale map."""
    plane = 1
    for d in x.shape[2:]:
        plane *= d
    return plane, x.shape[1], scale.shape[1]


def _laplace_qsample_map(x: Tensor, scale: Tensor, out: Tensor, noise: Optional[Tensor], u: Optional[Tensor],
                         noise_out: Optional[Tensor], x_mul: float, seed: int, offset: int) -> None:
    _cuda(x, scale, out, noise, u, noise_out)
    if x.numel() == 0:
        return
    plane, C, Cs = _map_dims(x, scale)
    check(_cabi.lib().ldiff_laplace_qsample_map(_ptr(x), _ptr(scale), _ptr(out), _ptr(noise), _ptr(u),
                                                _ptr(noise_out), x_mul, seed, offset, x.numel(), plane, C, Cs,
                                                _dt(x), _stream(x)))


def _scaled_residual(x: Tensor, eps: Tensor, scale: Tensor, out: Tensor, out_div: float) -> None:
    _cuda(x, eps, scale, out)
    if x.numel() == 0:
        return
    plane, C, Cs = _map_dims(x, scale)
    check(_cabi.lib().ldiff_scaled_residual(_ptr(x), _ptr(eps), _ptr(scale), _ptr(out), out_div, x.numel(),
                                            plane, C, Cs, _dt(x), _stream(x)))


def _plms_step(sample: Tensor, eps: Sequence[Tensor], mode: int, sample_coeff: float,
               alpha_diff: float, denom: float, out: Tensor) -> None:
    _cuda(sample, out, *eps)
    if sample.numel() == 0:
        return
    e = [_ptr(t) for t in eps] + [None] * (4 - len(eps))
    check(_cabi.lib().ldiff_plms_step(_ptr(sample), e[0], e[1], e[2], e[3], mode, sample_coeff,
                                      alpha_diff, denom, _ptr(out), sample.numel(), _dt(sample),
                                      _stream(sample)))


def _plms_step_noise(sample: Tensor, eps: Sequence[Tensor], mode: int, sample_coeff: float, alpha_diff: float,
                     denom: float, out: Tensor, clean: Tensor, noisy: Tensor, noise: Optional[Tensor],
                     u: Optional[Tensor], b: float, seed: int, offset: int) -> None:
    _cuda(sample, out, clean, noisy, noise, u, *eps)
    if sample.numel() == 0:
        return
    e = [_ptr(t) for t in eps] + [None] * (4 - len(eps))
    check(_cabi.lib().ldiff_plms_step_noise(_ptr(sample), e[0], e[1], e[2], e[3], mode, sample_coeff, alpha_diff,
                                            denom, _ptr(out), _ptr(clean), _ptr(noisy), _ptr(noise), _ptr(u), b,
                                            seed, offset, sample.numel(), _dt(sample), _stream(sample)))


def _decode_tail_gray(img: Tensor, rgb: Optional[Tensor], gray: Optional[Tensor]) -> None:
    _cuda(img, rgb, gray)
    if img.numel() == 0:
        return
    B, _, H, W = img.shape
    gstride = gray.stride(0) if gray is not None else 0
    check(_cabi.lib().ldiff_decode_tail_gray(_ptr(img), _ptr(rgb), _ptr(gray), B, H, W, gstride,
                                             _dt(img), _stream(img)))


def _decode_tail_fused(img: Tensor, rgb: Optional[Tensor], gray: Tensor, feat: Optional[Tensor], feat_channel: int,
                       small_rgb: Optional[Tensor], label: Optional[Tensor], label_plane: Optional[Tensor],
                       label_small: Optional[Tensor]) -> None:
    _cuda(img, rgb, gray, feat, small_rgb, label, label_plane, label_small)
    if img.numel() == 0:
        return
    B, _, H, W = img.shape
    check(_cabi.lib().ldiff_decode_tail_fused(
        _ptr(img), _ptr(rgb), _ptr(gray), B, H, W, gray.stride(0), _dt(img), _ptr(feat),
        _dt(feat) if feat is not None else _dt(img), feat.shape[1] if feat is not None else 1, feat_channel,
        _ptr(small_rgb), _ptr(label), _ptr(label_plane), label_plane.stride(0) if label_plane is not None else 0,
        _ptr(label_small), _stream(img)))


def _decode_tail_model_input(img: Tensor, rgb: Optional[Tensor], gray: Optional[Tensor], model_input: Tensor,
                             mean: Sequence[float], std: Sequence[float]) -> None:
    import ctypes
    _cuda(img, rgb, gray, model_input)
    if img.numel() == 0:
        return
    B, _, H, W = img.shape
    gstride = gray.stride(0) if gray is not None else 0
    m3 = (ctypes.c_float * 3)(*[float(v) for v in mean])
    s3 = (ctypes.c_float * 3)(*[float(v) for v in std])
    check(_cabi.lib().ldiff_decode_tail_model_input(_ptr(img), _ptr(rgb), _ptr(gray), _ptr(model_input), m3, s3,
                                                    B, H, W, gstride, _dt(img), _stream(img)))


def _bilinear_lift(src: Tensor, dst: Tensor, dst_channel: int, gray: bool) -> None:
    _cuda(src, dst)
    if src.numel() == 0 or dst.numel() == 0:
        return
    B, C, h, w = src.shape
    _, Ctot, H, W = dst.shape
    check(_cabi.lib().ldiff_bilinear_lift(_ptr(src), _dt(src), C, h, w, src.stride(0), src.stride(1),
                                          _ptr(dst), _dt(dst), Ctot, dst_channel, H, W, B,
                                          1 if gray else 0, _stream(src)))


def _bilinear_lift_multi(srcs: Sequence[Tensor], dst: Tensor, dst_channel: int, gray: bool) -> None:
    import ctypes
    _cuda(dst, *srcs)
    s0 = srcs[0]
    B, C, h, w = s0.shape
    _, Ctot, H, W = dst.shape
    ptrs = (ctypes.c_void_p * len(srcs))(*[t.data_ptr() for t in srcs])
    check(_cabi.lib().ldiff_bilinear_lift_multi(ptrs, len(srcs), _dt(s0), C, h, w, s0.stride(0), s0.stride(1),
                                                _ptr(dst), _dt(dst), Ctot, dst_channel, H, W, B,
                                                1 if gray else 0, _stream(s0)))


def _bilinear_lift_backward(grad_out: Tensor, dst_channel: int, grad_src: Tensor, gray: bool) -> None:
    _cuda(grad_out, grad_src)
    B, Ctot, H, W = grad_out.shape
    _, C, h, w = grad_src.shape
    check(_cabi.lib().ldiff_bilinear_lift_backward(_ptr(grad_out), Ctot, dst_channel, H, W, _ptr(grad_src), C, h, w,
                                                   grad_src.stride(0), grad_src.stride(1), B, 1 if gray else 0,
                                                   _stream(grad_out)))


def _head_logits(feat: Tensor, weight: Tensor, bias: Optional[Tensor], logits: Tensor,
                 clear: Optional[Tensor] = None) -> None:
    """``clear`` (optional int64 tensor): zeroed by the kernel as a side job (see ldiff.h)."""
    _cuda(feat, weight, bias, logits, clear)
    if feat.numel() == 0:
        if clear is not None:
            clear.zero_()
        return
    B, Cin = feat.shape[:2]
    hw = feat[0, 0].numel()
    check(_cabi.lib().ldiff_head_logits(_ptr(feat), _ptr(weight), _ptr(bias), _ptr(logits), B, Cin,
                                        weight.shape[0], hw, _dt(feat), _ptr(clear),
                                        0 if clear is None else clear.numel(), _stream(feat)))


def _lift_argmax(logits: Tensor, mask: Tensor) -> None:
    _cuda(logits, mask)
    if mask.numel() == 0:
        return
    B, K, h, w = logits.shape
    _, H, W = mask.shape
    check(_cabi.lib().ldiff_lift_argmax(_ptr(logits), _ptr(mask), B, K, h, w, H, W, _stream(logits)))


def _lift_argmax_hist(logits: Tensor, mask: Tensor, gt: Tensor, C: Tensor, status: Tensor, xchg=None,
                      channel: int = 0) -> None:
    """``xchg``: the C handle of a ``dist.ConfusionExchange`` (peer push as the kernel's tail) or None."""
    _cuda(logits, mask, gt, C, status)
    B, K, h, w = logits.shape
    _, H, W = mask.shape
    check(_cabi.lib().ldiff_lift_argmax_hist(_ptr(logits), _ptr(mask), _ptr(gt), _ptr(C), B, K, h, w, H, W, xchg,
                                             channel, _ptr(status), _stream(logits)))


def _ids16(inst: Tensor) -> bool:
    """Instance maps are int32 or — what Cellpose's ``eval`` returns below 65 536 labels (conductor.py:180) —
    uint16 (``torch.from_numpy`` of that array; an int16 view of the same bytes is taken as unsigned)."""
    if inst.dtype == torch.int32:
        return False
    if inst.dtype in (torch.uint16, torch.int16):
        return True
    raise TypeError(f"instance map must be int32 or uint16, got {inst.dtype}")


def _lut_paint_hist(inst: Tensor, lut: Tensor, mask: Tensor, gt: Tensor, C: Tensor, K: int, status: Tensor,
                    xchg=None, channel: int = 0) -> None:
    _cuda(inst, lut, mask, gt, C, status)
    B = inst.shape[0]
    n = inst[0].numel()
    lut_stride = lut.stride(0) if lut.dim() == 2 else 0
    fn = _cabi.lib().ldiff_lut_paint_hist_u16 if _ids16(inst) else _cabi.lib().ldiff_lut_paint_hist
    check(fn(_ptr(inst), _ptr(lut), _ptr(mask), _ptr(gt), _ptr(C), n, B, lut.shape[-1],
                                           lut_stride, K, xchg, channel, _ptr(status), _stream(inst)))


def _cell_classify(feats: Tensor, weight: Tensor, bias: Optional[Tensor], inst_ids: Tensor,
                   lut: Tensor, logits_out: Optional[Tensor], status: Tensor, clear: Optional[Tensor] = None) -> None:
    """feats [B,N,Cin] (or [N,Cin]); lut [B,lut_size] (or [lut_size]); ``clear``: see ``_head_logits``."""
    _cuda(feats, weight, bias, inst_ids, lut, logits_out, status, clear)
    if feats.numel() == 0:                                # no instances: the LUT keeps its background entries
        if clear is not None:
            clear.zero_()
        return
    B = feats.shape[0] if feats.dim() == 3 else 1
    N, Cin = feats.shape[-2], feats.shape[-1]
    lut_stride = lut.stride(0) if lut.dim() == 2 else 0
    check(_cabi.lib().ldiff_cell_classify(_ptr(feats), _ptr(weight), _ptr(bias), _ptr(inst_ids),
                                          _ptr(lut), lut.shape[-1], lut_stride, _ptr(logits_out), N, B, Cin,
                                          weight.shape[0], _dt(feats), _ptr(clear),
                                          0 if clear is None else clear.numel(), _ptr(status), _stream(feats)))


def _copy_planes_u8(src: Tensor, dst: Tensor) -> None:
    _cuda(src, dst)
    check(_cabi.lib().ldiff_copy_planes_u8(_ptr(src), _ptr(dst), src[0].numel(), src.shape[0], dst.stride(0),
                                           _stream(src)))


def _lut_paint(inst: Tensor, lut: Tensor, mask: Tensor, status: Tensor) -> None:
    _cuda(inst, lut, mask, status)
    if inst.numel() == 0:
        return
    B = inst.shape[0]
    n = inst[0].numel()
    lut_stride = lut.stride(0) if lut.dim() == 2 else 0
    fn = _cabi.lib().ldiff_lut_paint_u16 if _ids16(inst) else _cabi.lib().ldiff_lut_paint
    check(fn(_ptr(inst), _ptr(lut), _ptr(mask), n, B, lut.shape[-1], lut_stride, _ptr(status), _stream(inst)))


def _argmax_channels(x: Tensor, out: Tensor) -> None:
    _cuda(x, out)
    if x.numel() == 0:
        return
    B, K = x.shape[:2]
    check(_cabi.lib().ldiff_argmax_channels(_ptr(x), _ptr(out), B, K, x[0, 0].numel(), _dt(x),
                                            _stream(x)))


def _confusion_hist(pred: Tensor, gt: Tensor, gt_lut: Optional[Tensor], C: Tensor, K: int,
                    status: Tensor) -> None:
    _cuda(pred, gt, gt_lut, C, status)
    if pred.numel() == 0:
        return
    check(_cabi.lib().ldiff_confusion_hist(_ptr(pred), _ptr(gt), _ptr(gt_lut), _ptr(C), pred.numel(),
                                           K, _ptr(status), _stream(pred)))


def _confusion_hist_batched(pred: Tensor, gt: Tensor, gt_lut: Optional[Tensor], C: Tensor, K: int,
                            status: Tensor) -> None:
    _cuda(pred, gt, gt_lut, C, status)
    if pred.numel() == 0:
        return
    n_images = pred.shape[0]
    check(_cabi.lib().ldiff_confusion_hist_batched(_ptr(pred), _ptr(gt), _ptr(gt_lut), _ptr(C),
                                                   pred[0].numel(), n_images, K, _ptr(status), _stream(pred)))


def _labels_to_u8(x: Tensor, out: Tensor) -> None:
    _cuda(x, out)
    check(_cabi.lib().ldiff_labels_to_u8(_ptr(x), _ptr(out), x.numel(), _stream(x)))


def _sw_accumulate(pred: Tensor, gauss: Optional[Tensor], acc: Tensor, npred: Tensor, y0: int, x0: int) -> None:
    _cuda(pred, gauss, acc, npred)
    K, th, tw = pred.shape
    _, H, W = acc.shape
    check(_cabi.lib().ldiff_sw_accumulate(_ptr(pred), _ptr(gauss), _ptr(acc), _ptr(npred), K, th, tw, H, W, y0, x0,
                                          _stream(pred)))


def _sw_tta_merge(preds: Sequence[Tensor], flips: Sequence[int], out: Tensor) -> None:
    import ctypes
    _cuda(out, *preds)
    K, th, tw = out.shape
    n = len(preds)
    ptrs = (ctypes.c_void_p * n)(*[t.data_ptr() for t in preds])
    fl = (ctypes.c_int * n)(*[int(f) for f in flips])
    check(_cabi.lib().ldiff_sw_tta_merge(ptrs, fl, n, _ptr(out), K, th, tw, _stream(out)))


def _sw_finalize_argmax(acc: Tensor, npred: Tensor, seg: Tensor, logits_out: Optional[Tensor],
                        status: Tensor) -> None:
    _cuda(acc, npred, seg, logits_out, status)
    check(_cabi.lib().ldiff_sw_finalize_argmax(_ptr(acc), _ptr(npred), _ptr(seg), _ptr(logits_out), acc.shape[0],
                                               npred.numel(), _ptr(status), _stream(acc)))


def _infonce_sample(labels: Tensor, n_neg: int, cap: int, seed: int, offset: int, pb: Tensor, pa: Tensor, pq: Tensor,
                    neg: Tensor, n_valid: Tensor, status: Tensor) -> None:
    _cuda(labels, pb, pa, pq, neg, n_valid, status)
    B = labels.shape[0]
    check(_cabi.lib().ldiff_infonce_sample(_ptr(labels), B, labels[0].numel(), n_neg, cap, seed & (2 ** 64 - 1),
                                           offset & (2 ** 64 - 1), _ptr(pb), _ptr(pa), _ptr(pq), _ptr(neg),
                                           _ptr(n_valid), _ptr(status), _stream(labels)))


def _infonce_forward(feat: Tensor, pb: Tensor, pa: Tensor, pq: Tensor, neg: Tensor, loss: Tensor, lse: Tensor,
                     temperature: float) -> None:
    _cuda(feat, pb, pa, pq, neg, loss, lse)
    B, C = feat.shape[:2]
    check(_cabi.lib().ldiff_infonce_forward(_ptr(feat), _ptr(pb), _ptr(pa), _ptr(pq), _ptr(neg), _ptr(loss),
                                            _ptr(lse), C, feat[0, 0].numel(), neg.shape[1], pa.numel(),
                                            temperature, _stream(feat)))


def _infonce_backward(feat: Tensor, pb: Tensor, pa: Tensor, pq: Tensor, neg: Tensor, lse: Tensor, gscale: Tensor,
                      grad: Tensor, temperature: float) -> None:
    _cuda(feat, pb, pa, pq, neg, lse, gscale, grad)
    B, C = feat.shape[:2]
    check(_cabi.lib().ldiff_infonce_backward(_ptr(feat), _ptr(pb), _ptr(pa), _ptr(pq), _ptr(neg), _ptr(lse),
                                             _ptr(gscale), _ptr(grad), C, feat[0, 0].numel(), neg.shape[1],
                                             pa.numel(), temperature, _stream(feat)))


torch.library.custom_op("ldiff::infonce_sample", mutates_args=("pb", "pa", "pq", "neg", "n_valid", "status"))(_infonce_sample)
torch.library.custom_op("ldiff::infonce_forward", mutates_args=("loss", "lse"))(_infonce_forward)
torch.library.custom_op("ldiff::infonce_backward", mutates_args=("grad",))(_infonce_backward)
torch.library.custom_op("ldiff::sw_accumulate", mutates_args=("acc", "npred"))(_sw_accumulate)
torch.library.custom_op("ldiff::sw_tta_merge", mutates_args=("out",))(_sw_tta_merge)
torch.library.custom_op("ldiff::sw_finalize_argmax", mutates_args=("seg", "logits_out", "status"))(_sw_finalize_argmax)
torch.library.custom_op("ldiff::laplace_qsample", mutates_args=("out", "noise_out"))(_laplace_qsample)
torch.library.custom_op("ldiff::plms_step", mutates_args=("out",))(_plms_step)
torch.library.custom_op("ldiff::plms_step_noise", mutates_args=("out", "noisy"))(_plms_step_noise)
torch.library.custom_op("ldiff::laplace_qsample_map", mutates_args=("out", "noise_out"))(_laplace_qsample_map)
torch.library.custom_op("ldiff::scaled_residual", mutates_args=("out",))(_scaled_residual)
torch.library.custom_op("ldiff::decode_tail_gray", mutates_args=("rgb", "gray"))(_decode_tail_gray)
torch.library.custom_op("ldiff::decode_tail_fused", mutates_args=("rgb", "gray", "feat", "small_rgb", "label_plane", "label_small"))(_decode_tail_fused)
torch.library.custom_op("ldiff::decode_tail_model_input",
                        mutates_args=("rgb", "gray", "model_input"))(_decode_tail_model_input)
torch.library.custom_op("ldiff::bilinear_lift", mutates_args=("dst",))(_bilinear_lift)
torch.library.custom_op("ldiff::bilinear_lift_multi", mutates_args=("dst",))(_bilinear_lift_multi)
torch.library.custom_op("ldiff::bilinear_lift_backward", mutates_args=("grad_src",))(_bilinear_lift_backward)
torch.library.custom_op("ldiff::head_logits", mutates_args=("logits", "clear"))(_head_logits)
torch.library.custom_op("ldiff::lift_argmax", mutates_args=("mask",))(_lift_argmax)
# (the *_hist entry points take an opaque exchange handle and are not dispatcher ops; use ops.lift_argmax_hist / ops.lut_paint_hist)
torch.library.custom_op("ldiff::cell_classify", mutates_args=("lut", "logits_out", "status", "clear"))(_cell_classify)
torch.library.custom_op("ldiff::copy_planes_u8", mutates_args=("dst",))(_copy_planes_u8)
torch.library.custom_op("ldiff::lut_paint", mutates_args=("mask", "status"))(_lut_paint)
torch.library.custom_op("ldiff::argmax_channels", mutates_args=("out",))(_argmax_channels)
torch.library.custom_op("ldiff::confusion_hist", mutates_args=("C", "status"))(_confusion_hist)
torch.library.custom_op("ldiff::confusion_hist_batched", mutates_args=("C", "status"))(_confusion_hist_batched)
torch.library.custom_op("ldiff::labels_to_u8", mutates_args=("out",))(_labels_to_u8)


# ---------------------------------------------------------------------------
# functional wrappers
# ---------------------------------------------------------------------------

def laplace_qsample(x: Tensor, b: float, *, noise: Optional[Tensor] = None, u: Optional[Tensor] = None,
                    seed: int = 0, offset: int = 0, return_noise: bool = False, out: Optional[Tensor] = None):
    """noisy = x + Laplace(0, b) noise  (ldiffusion.py:234-237).

    ``noise``: injected noise tensor (parity mode);  ``u``: injected uniforms in
    (-1, 1);  otherwise Philox(seed, offset).  Element i of the flattened tensor
    consumes word i%4 of Philox counter ``offset + i//4``: advance ``offset`` by
    ``ceil(numel/4)`` between calls to continue the stream.
    """
    _dense(x, "x")
    if noise is not None and u is not None:
        raise ValueError("pass at most one of noise / u")
    for t in (noise, u):
        if t is not None and (t.shape != x.shape or t.dtype != x.dtype or not t.is_contiguous()):
            raise ValueError("injected tensor must match x in shape, dtype and be contiguous")
    out = torch.empty_like(x) if out is None else out
    nz = torch.empty_like(x) if return_noise else None
    _laplace_qsample(x, out, noise, u, nz, float(b), int(seed), int(offset))
    return (out, nz) if return_noise else out


def _check_scale_map(x: Tensor, scale: Tensor):
    if x.dim() < 3:
        raise ValueError("x must be [B,C,...]")
    if scale.dtype != x.dtype or not scale.is_contiguous() or scale.dim() != x.dim() \
            or scale.shape[0] != x.shape[0] or scale.shape[2:] != x.shape[2:] \
            or scale.shape[1] not in (1, x.shape[1]):
        raise ValueError("scale must be contiguous [B,C,...] or [B,1,...] matching x in dtype and size")


def laplace_qsample_map(x: Tensor, scale: Tensor, *, noise: Optional[Tensor] = None, u: Optional[Tensor] = None,
                        seed: int = 0, offset: int = 0, x_mul: float = 1.0, return_noise: bool = False,
                        out: Optional[Tensor] = None):
    """noisy = x_mul * x + Laplace(0, 1) noise * scale  (segmentor.py:339,344-345; the multimodal variant).

    ``scale`` is the per-pixel map (``depth_resized``): [B,C,h,w], or [B,1,h,w] broadcast over the
    channels instead of the reference's ``.repeat(1, C, 1, 1)`` copy.  Randomness as in
    ``laplace_qsample``; ``return_noise`` gives the unit noise (the reference's ``noise`` tensor)."""
    _dense(x, "x")
    _check_scale_map(x, scale)
    if noise is not None and u is not None:
        raise ValueError("pass at most one of noise / u")
    for t in (noise, u):
        if t is not None and (t.shape != x.shape or t.dtype != x.dtype or not t.is_contiguous()):
            raise ValueError("injected tensor must match x in shape, dtype and be contiguous")
    out = torch.empty_like(x) if out is None else out
    nz = torch.empty_like(x) if return_noise else None
    if x.numel() == 0:
        return (out, nz) if return_noise else out
    _laplace_qsample_map(x, scale, out, noise, u, nz, float(x_mul), int(seed), int(offset))
    return (out, nz) if return_noise else out


def scaled_residual(x: Tensor, eps: Tensor, scale: Tensor, *, out_div: float = 1.0,
                    out: Optional[Tensor] = None) -> Tensor:
    """(x - eps * scale) / out_div  (segmentor.py:375,379: the inverse of ``laplace_qsample_map`` with
    the UNet's noise prediction, then the VAE's 1/0.18215)."""
    _dense(x, "x")
    _check_scale_map(x, scale)
    if eps.shape != x.shape or eps.dtype != x.dtype or not eps.is_contiguous():
        raise ValueError("eps must match x in shape, dtype and be contiguous")
    if not out_div:
        raise ValueError("out_div must be non-zero")
    out = torch.empty_like(x) if out is None else out
    if x.numel() == 0:
        return out
    _scaled_residual(x, eps, scale, out, float(out_div))
    return out


def plms_step(sample: Tensor, eps: Sequence[Tensor], mode: int, sample_coeff: float, alpha_diff: float,
              denom: float, out: Optional[Tensor] = None) -> Tensor:
    """prev = sample_coeff*sample - (alpha_diff*eps_hat)/denom; eps[0] is the newest output."""
    need = {0: 1, 1: 2, 2: 2, 3: 3, 4: 4}[mode]
    if len(eps) < need:
        raise ValueError(f"mode {mode} needs {need} model outputs")
    eps = list(eps[:need])
    _dense(sample, "sample")
    for e in eps:
        if e.shape != sample.shape or e.dtype != sample.dtype or not e.is_contiguous():
            raise ValueError("model outputs must match the sample in shape, dtype and be contiguous")
    out = torch.empty_like(sample) if out is None else out
    _plms_step(sample, eps, mode, float(sample_coeff), float(alpha_diff), float(denom), out)
    return out


def plms_step_noise(sample: Tensor, eps: Sequence[Tensor], mode: int, sample_coeff: float, alpha_diff: float,
                    denom: float, clean: Tensor, b: float, *, noise: Optional[Tensor] = None,
                    u: Optional[Tensor] = None, seed: int = 0, offset: int = 0, out: Optional[Tensor] = None,
                    noisy_out: Optional[Tensor] = None):
    """``plms_step`` on (sample, eps) and ``laplace_qsample`` on ``clean`` in ONE launch; returns
    (prev_sample, noisy).  Bit-identical to the two separate operators."""
    need = {0: 1, 1: 2, 2: 2, 3: 3, 4: 4}[mode]
    if len(eps) < need:
        raise ValueError(f"mode {mode} needs {need} model outputs")
    eps = list(eps[:need])
    _dense(sample, "sample")
    for e in eps + [clean] + [t for t in (noise, u) if t is not None]:
        if e.shape != sample.shape or e.dtype != sample.dtype or not e.is_contiguous():
            raise ValueError("all tensors must match the sample in shape, dtype and be contiguous")
    if noise is not None and u is not None:
        raise ValueError("pass at most one of noise / u")
    out = torch.empty_like(sample) if out is None else out
    noisy_out = torch.empty_like(sample) if noisy_out is None else noisy_out
    _plms_step_noise(sample, eps, mode, float(sample_coeff), float(alpha_diff), float(denom), out, clean, noisy_out,
                     noise, u, float(b), int(seed), int(offset))
    return out, noisy_out


IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def decode_tail_model_input(img: Tensor, *, mean=IMAGENET_MEAN, std=IMAGENET_STD, want_rgb: bool = True,
                            want_gray: bool = False, gray_out: Optional[Tensor] = None):
    """Decode tail with the segmentor's model input fused in: returns (uint8 RGB | None, uint8 gray |
    None, fp32 [B,3,H,W] = Normalize(mean, std)(ToTensor(image))) — the hand-off the reference does
    through PIL + torchvision at segmentor.py:107-108 / :533-534, without leaving the device."""
    if img.dim() != 4 or img.shape[1] != 3:
        raise ValueError("img must be [B,3,H,W]")
    _cuda(img)
    _dense(img, "img")
    B, _, H, W = img.shape
    rgb = torch.empty((B, H, W, 3), dtype=torch.uint8, device=img.device) if want_rgb else None
    gray = gray_out
    if gray is None and want_gray:
        gray = torch.empty((B, H, W), dtype=torch.uint8, device=img.device)
    mi = torch.empty((B, 3, H, W), dtype=torch.float32, device=img.device)
    _decode_tail_model_input(img, rgb, gray, mi, list(mean), list(std))
    return rgb, gray, mi


def decode_tail_gray(img: Tensor, *, want_rgb: bool = True, gray_out: Optional[Tensor] = None,
                     want_gray: bool = True, rgb_out: Optional[Tensor] = None):
    """[B,3,H,W] decoder output -> (uint8 RGB [B,H,W,3] | None, uint8 gray [B,H,W] | None).

    ``gray_out`` may be a [B,H,W] view whose planes are dense (e.g. slot i of a
    [B,n+1,H,W] pixel-vector tensor), so the per-step concat costs nothing.
    """
    if img.dim() != 4 or img.shape[1] != 3:
        raise ValueError("img must be [B,3,H,W]")
    _dense(img, "img")
    B, _, H, W = img.shape
    rgb = rgb_out
    if rgb is None and want_rgb:
        rgb = torch.empty((B, H, W, 3), dtype=torch.uint8, device=img.device)
    gray = gray_out
    if gray is None and want_gray:
        gray = torch.empty((B, H, W), dtype=torch.uint8, device=img.device)
    if gray is not None:
        if gray.shape != (B, H, W) or gray.dtype != torch.uint8 or (H * W and gray.stride()[1:] != (W, 1)):
            raise ValueError("gray_out must be uint8 [B,H,W] with dense planes")
    if rgb is None and gray is None:
        raise ValueError("nothing to compute")
    _decode_tail_gray(img, rgb, gray)
    return rgb, gray


def decode_tail_fused(img: Tensor, gray_out: Tensor, *, rgb_out: Optional[Tensor] = None,
                      feat_out: Optional[Tensor] = None, feat_channel: int = 0, small_rgb_out: Optional[Tensor] = None,
                      label: Optional[Tensor] = None, label_plane_out: Optional[Tensor] = None,
                      label_small_out: Optional[Tensor] = None) -> None:
    """``decode_tail_gray`` with the per-step consumers of the decoder output fused into the same pass:
    channel ``feat_channel`` of ``feat_out`` [B,n,H/16,W/16] = weighted gray of the 16x bilinear
    down-sample (ldiffusion.py:240-247), ``small_rgb_out`` [B,3,H/16,W/16] = that down-sample itself,
    ``label_plane_out`` = ``label`` copied into a [B,H,W] slot of the pixel vectors,
    ``label_small_out`` [B,1,H/16,W/16] = the label's bilinear down-sample truncated to uint8
    (ldiffusion.py:224-226).  H and W must be multiples of 16 (otherwise use the separate operators)."""
    if img.dim() != 4 or img.shape[1] != 3:
        raise ValueError("img must be [B,3,H,W]")
    _dense(img, "img")
    B, _, H, W = img.shape
    if H % 16 or W % 16:
        raise ValueError("decode_tail_fused needs H and W to be multiples of 16")
    fh, fw = H // 16, W // 16
    if gray_out.shape != (B, H, W) or gray_out.dtype != torch.uint8 or gray_out.stride()[1:] != (W, 1):
        raise ValueError("gray_out must be uint8 [B,H,W] with dense planes")
    if feat_out is not None:
        _dense(feat_out, "feat_out")
        if feat_out.dim() != 4 or feat_out.shape[0] != B or feat_out.shape[2:] != (fh, fw) \
                or feat_out.dtype not in (img.dtype, torch.float32) or not 0 <= feat_channel < feat_out.shape[1]:
            raise ValueError("feat_out must be [B,n,H/16,W/16] in the image dtype or fp32")
    if small_rgb_out is not None:
        _dense(small_rgb_out, "small_rgb_out")
        if small_rgb_out.shape != (B, 3, fh, fw) or small_rgb_out.dtype != img.dtype:
            raise ValueError("small_rgb_out must be [B,3,H/16,W/16] in the image dtype")
    if label_plane_out is not None or label_small_out is not None:
        if label is None or label.dtype != torch.uint8 or label.shape != (B, H, W):
            raise ValueError("label must be uint8 [B,H,W]")
        _dense(label, "label")
    if label_plane_out is not None and (label_plane_out.shape != (B, H, W) or label_plane_out.dtype != torch.uint8
                                        or label_plane_out.stride()[1:] != (W, 1)):
        raise ValueError("label_plane_out must be uint8 [B,H,W] with dense planes")
    if label_small_out is not None:
        _dense(label_small_out, "label_small_out")
        if label_small_out.numel() != B * fh * fw or label_small_out.dtype != torch.uint8:
            raise ValueError("label_small_out must be uint8 [B,1,H/16,W/16]")
    _decode_tail_fused(img, rgb_out, gray_out, feat_out, int(feat_channel), small_rgb_out, label, label_plane_out,
                       label_small_out)


def bilinear_lift(src: Tensor, size, *, out: Optional[Tensor] = None, out_channel: int = 0,
                  gray: bool = False, out_dtype=None) -> Tensor:
    """F.interpolate(src, size, mode='bilinear', align_corners=False) [+ weighted
    gray], written into channels of ``out`` ([B,Ctot,H,W]) starting at ``out_channel``."""
    if src.dim() != 4:
        raise ValueError("src must be [B,C,h,w]")
    if src.stride(3) != 1 or src.stride(2) != src.shape[3]:
        raise ValueError("src planes must be dense")
    B, C = src.shape[:2]
    H, W = size
    if out is None:
        out = torch.empty((B, 1 if gray else C, H, W), dtype=out_dtype or src.dtype, device=src.device)
    _dense(out, "out")
    if out.shape[0] != B or out.shape[2:] != (H, W):
        raise ValueError("out must be [B,Ctot,H,W]")
    _bilinear_lift(src, out, out_channel, gray)
    return out


def bilinear_lift_backward(grad_out: Tensor, src_shape, *, out_channel: int = 0, gray: bool = False) -> Tensor:
    """Adjoint of ``bilinear_lift`` (fp32): gradient w.r.t. a ``src_shape`` = [B,C,h,w] source of the lift whose
    result sits in channels ``out_channel ...`` of ``grad_out`` [B,Ctot,H,W].  Round-2 widening (training
    caller, ldiffusion.py:240-252); not measured yet."""
    if grad_out.dtype != torch.float32 or grad_out.dim() != 4:
        raise TypeError("grad_out must be fp32 [B,Ctot,H,W]")
    _cuda(grad_out)
    _dense(grad_out, "grad_out")
    B, C, h, w = src_shape
    if grad_out.shape[0] != B or (gray and C != 3):
        raise ValueError("grad_out / src_shape mismatch")
    grad_src = torch.zeros((B, C, h, w), dtype=torch.float32, device=grad_out.device)
    if grad_src.numel() and grad_out.numel():
        _bilinear_lift_backward(grad_out, int(out_channel), grad_src, bool(gray))
    return grad_src


class _LiftFn(torch.autograd.Function):
    """bilinear_lift (+ gray) with a backward: the lift kernels run under no_grad everywhere else."""

    @staticmethod
    def forward(ctx, src, size, gray):
        ctx.src_shape, ctx.gray, ctx.src_dtype = tuple(src.shape), bool(gray), src.dtype
        return bilinear_lift(src.detach().contiguous(), size, gray=gray, out_dtype=torch.float32)

    @staticmethod
    def backward(ctx, grad):
        g = bilinear_lift_backward(grad.float().contiguous(), ctx.src_shape, gray=ctx.gray)
        return g.to(ctx.src_dtype), None, None


def bilinear_lift_autograd(src: Tensor, size, *, gray: bool = False) -> Tensor:
    """Differentiable ``bilinear_lift`` (fp32 result): forward = the lift kernel, backward = its adjoint kernel."""
    return _LiftFn.apply(src, tuple(size), gray)


def bilinear_lift_multi(srcs: Sequence[Tensor], size, *, out: Optional[Tensor] = None, out_channel: int = 0,
                        gray: bool = False, out_dtype=None) -> Tensor:
    """lift (+ gray) of up to 8 same-shaped sources in one launch; source i is written at channel
    ``out_channel + i * (1 if gray else C)`` of ``out`` — lift + gray + torch.cat of ldiffusion.py:240-247."""
    s0 = srcs[0]
    for t in srcs:
        if t.shape != s0.shape or t.dtype != s0.dtype or t.stride() != s0.stride():
            raise ValueError("sources must share shape, dtype and strides")
        if t.stride(3) != 1 or t.stride(2) != t.shape[3]:
            raise ValueError("source planes must be dense")
    if len(srcs) > 8:
        raise ValueError("at most 8 sources per launch")
    B, C = s0.shape[:2]
    H, W = size
    per = 1 if gray else C
    if out is None:
        out = torch.empty((B, per * len(srcs), H, W), dtype=out_dtype or s0.dtype, device=s0.device)
    _dense(out, "out")
    _bilinear_lift_multi(list(srcs), out, out_channel, gray)
    return out


def head_logits(feat: Tensor, weight: Tensor, bias: Optional[Tensor] = None) -> Tensor:
    """1x1 conv Cin->K (conductor.py:127): feat [B,Cin,h,w] -> fp32 logits [B,K,h,w]."""
    _dense(feat, "feat"); _dense(weight, "weight")
    if weight.dtype != feat.dtype:
        raise ValueError("weight dtype must match feat")
    B, Cin, h, w = feat.shape
    K = weight.shape[0]
    if weight.shape[1] != Cin:
        raise ValueError("weight must be [K,Cin]")
    bias = None if bias is None else bias.float().contiguous()
    logits = torch.empty((B, K, h, w), dtype=torch.float32, device=feat.device)
    _head_logits(feat, weight.reshape(K, Cin), bias, logits)
    return logits


def lift_argmax(logits: Tensor, size) -> Tensor:
    """argmax(softmax(F.interpolate(logits, size)), 1) as uint8 [B,H,W]
    (conductor.py:135 + segmentor.py:536) without materialising the lift."""
    _dense(logits, "logits")
    if logits.dtype != torch.float32:
        raise TypeError("logits must be fp32")
    mask = torch.empty((logits.shape[0], size[0], size[1]), dtype=torch.uint8, device=logits.device)
    _lift_argmax(logits, mask)
    return mask


def head_argmax(feat: Tensor, weight: Tensor, bias: Optional[Tensor], size, return_logits: bool = False):
    logits = head_logits(feat, weight, bias)
    mask = lift_argmax(logits, size)
    return (mask, logits) if return_logits else mask


def cell_classify(inst_feats: Tensor, weight: Tensor, bias: Optional[Tensor], inst_ids: Tensor,
                  lut_size: int, *, lut: Optional[Tensor] = None, return_logits: bool = False):
    """conductor.py:218-221 -> class LUT.  inst_feats [N,Cin] -> uint8 [lut_size], or a batch
    [B,N,Cin] -> [B,lut_size] in one launch (``inst_ids`` [N] shared); lut[..., 0] = background."""
    _cuda(inst_feats, weight, inst_ids)
    _dense(inst_feats, "inst_feats"); _dense(weight, "weight")
    if inst_feats.shape[-1] % 8:
        raise ValueError("feature width must be a multiple of 8")
    if inst_ids.dtype != torch.int32:
        inst_ids = inst_ids.to(torch.int32)
    batched = inst_feats.dim() == 3
    if lut is None:
        shape = (inst_feats.shape[0], lut_size) if batched else (lut_size,)
        lut = torch.zeros(shape, dtype=torch.uint8, device=inst_feats.device)
    K = weight.shape[0]
    lo = torch.empty(tuple(inst_feats.shape[:-1]) + (K,), dtype=torch.float32,
                     device=inst_feats.device) if return_logits else None
    bias = None if bias is None else bias.float().contiguous()
    _cell_classify(inst_feats, weight, bias, inst_ids.contiguous(), lut, lo, status_word(inst_feats.device))
    return (lut, lo) if return_logits else lut


def copy_planes_u8(src: Tensor, dst: Tensor) -> Tensor:
    """dst[b] = src[b] for uint8 [B,H,W] planes, dst possibly a strided slot (dense planes)."""
    _cuda(src, dst)
    if src.dtype != torch.uint8 or dst.dtype != torch.uint8 or src.shape != dst.shape:
        raise ValueError("uint8 tensors of the same shape expected")
    _dense(src, "src")
    n = src[0].numel()
    if n % 16 or dst.stride(0) % 16 or dst[0].stride() != src[0].stride():
        dst.copy_(src)                                   # odd sizes: plain strided copy
        return dst
    _copy_planes_u8(src, dst)
    return dst


def lut_paint(inst: Tensor, lut: Tensor, out: Optional[Tensor] = None) -> Tensor:
    """mask[b,y,x] = lut[b][inst[b,y,x]] (conductor.py:224-231 + segmentor.py:536); ``inst`` int32 or uint16."""
    _ids16(inst)
    _cuda(inst, lut)
    if inst.dim() == 2:
        inst = inst.unsqueeze(0)
    _dense(inst, "inst"); _dense(lut, "lut")
    if lut.dim() == 2 and lut.shape[0] != inst.shape[0]:
        raise ValueError("per-image LUTs must be [B,lut_size]")
    mask = torch.empty(inst.shape, dtype=torch.uint8, device=inst.device) if out is None else out
    _lut_paint(inst, lut, mask, status_word(inst.device))
    return mask


def _check_hist_args(mask_shape, gt: Tensor, C: Optional[Tensor], K: int, device):
    if gt.dtype != torch.uint8 or tuple(gt.shape) != tuple(mask_shape):
        raise ValueError("gt must be a uint8 map of the mask's shape")
    _dense(gt, "gt")
    if C is None:
        C = torch.zeros((K + 1, K), dtype=torch.int64, device=device)
    elif C.shape != (K + 1, K) or C.dtype != torch.int64 or not C.is_contiguous():
        raise ValueError("out must be a contiguous int64 [(K+1),K] tensor")
    return C


def lift_argmax_hist(logits: Tensor, size, gt: Tensor, *, out: Optional[Tensor] = None,
                     mask_out: Optional[Tensor] = None, exchange=None, channel: int = 0):
    """``lift_argmax`` and ``confusion_hist(mask, gt)`` in ONE kernel: returns (mask, C); C accumulates
    over the whole batch.  ``exchange``: a ``dist.ConfusionExchange`` whose peer push rides as the
    kernel's tail.  Shapes the fused kernel does not take (K > 15, lifts below 4x) run the two
    separate kernels."""
    _dense(logits, "logits")
    if logits.dtype != torch.float32:
        raise TypeError("logits must be fp32")
    B, K, h, w = logits.shape
    H, W = size
    mask = torch.empty((B, H, W), dtype=torch.uint8, device=logits.device) if mask_out is None else mask_out
    C = _check_hist_args(mask.shape, gt, out, K, logits.device)
    st = status_word(logits.device)
    try:
        _lift_argmax_hist(logits, mask, gt, C, st, None if exchange is None else exchange._h, channel)
    except LdiffError as e:
        if "unsupported" not in str(e):
            raise
        _lift_argmax(logits, mask)
        if exchange is None:
            _confusion_hist(mask.view(-1), gt.view(-1), None, C, K, st)
        else:
            exchange.hist_push(mask.view(-1), gt.view(-1), C, channel=channel)
    return mask, C


def lut_paint_hist(inst: Tensor, lut: Tensor, gt: Tensor, num_classes: int, *, out: Optional[Tensor] = None,
                   mask_out: Optional[Tensor] = None, exchange=None, channel: int = 0):
    """``lut_paint`` and ``confusion_hist(mask, gt)`` in ONE kernel (6 B/pixel instead of 5 + 2): returns
    (mask, C); ``inst`` int32 or uint16 (4 B/pixel).  Falls back to the two separate kernels for K > 15 or planes
    that are not 16-pixel aligned."""
    _ids16(inst)
    _cuda(inst, lut, gt)
    if inst.dim() == 2:
        inst, gt = inst.unsqueeze(0), gt.unsqueeze(0) if gt.dim() == 2 else gt
    _dense(inst, "inst"); _dense(lut, "lut")
    if lut.dim() == 2 and lut.shape[0] != inst.shape[0]:
        raise ValueError("per-image LUTs must be [B,lut_size]")
    K = int(num_classes)
    mask = torch.empty(inst.shape, dtype=torch.uint8, device=inst.device) if mask_out is None else mask_out
    C = _check_hist_args(mask.shape, gt, out, K, inst.device)
    st = status_word(inst.device)
    try:
        _lut_paint_hist(inst, lut, mask, gt, C, K, st, None if exchange is None else exchange._h, channel)
    except LdiffError as e:
        if "unsupported" not in str(e) and "aligned" not in str(e):
            raise
        _lut_paint(inst, lut, mask, st)
        if exchange is None:
            _confusion_hist(mask.view(-1), gt.view(-1), None, C, K, st)
        else:
            exchange.hist_push(mask.view(-1), gt.view(-1), C, channel=channel)
    return mask, C


def argmax_channels(x: Tensor) -> Tensor:
    """torch.argmax(x, dim=1) (first maximum) as uint8, for [B,K,...] inputs."""
    _dense(x, "x")
    if x.shape[1] > 255:
        raise ValueError("at most 255 classes")
    out = torch.empty((x.shape[0],) + tuple(x.shape[2:]), dtype=torch.uint8, device=x.device)
    _argmax_channels(x, out)
    return out


def labels_to_u8(x: Tensor) -> Tensor:
    if x.dtype == torch.uint8:
        return x.contiguous()
    if x.dtype != torch.int64:
        x = x.to(torch.int64)
    x = x.contiguous()
    out = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    _labels_to_u8(x, out)
    return out


def confusion_hist(pred: Tensor, gt: Tensor, num_classes: int, *, out: Optional[Tensor] = None,
                   gt_lut: Optional[Tensor] = None) -> Tensor:
    """Accumulates C[(K+1),K] (int64) += histogram of (gt, pred) pairs; uint8 inputs."""
    if pred.dtype != torch.uint8 or gt.dtype != torch.uint8:
        raise TypeError("pred and gt must be uint8 label maps")
    _cuda(pred, gt)
    if pred.numel() != gt.numel():
        raise ValueError("pred and gt must have the same number of pixels")
    _dense(pred, "pred"); _dense(gt, "gt")
    K = int(num_classes)
    if out is None:
        out = torch.zeros((K + 1, K), dtype=torch.int64, device=pred.device)
    elif out.shape != (K + 1, K) or out.dtype != torch.int64 or not out.is_contiguous():
        raise ValueError("out must be a contiguous int64 [(K+1),K] tensor")
    if gt_lut is not None and (gt_lut.dtype != torch.uint8 or gt_lut.numel() != 256):
        raise ValueError("gt_lut must be 256 uint8 entries")
    _confusion_hist(pred, gt, gt_lut, out, K, status_word(pred.device))
    return out


def confusion_hist_batched(pred: Tensor, gt: Tensor, num_classes: int, *, out: Optional[Tensor] = None,
                           gt_lut: Optional[Tensor] = None) -> Tensor:
    """Per-image matrices in ONE launch: uint8 [N,...] x2 -> int64 [N,(K+1),K] (accumulates)."""
    if pred.dtype != torch.uint8 or gt.dtype != torch.uint8:
        raise TypeError("pred and gt must be uint8 label maps")
    _cuda(pred, gt)
    if pred.shape != gt.shape or pred.dim() < 2:
        raise ValueError("pred and gt must be [N,...] of the same shape")
    _dense(pred, "pred"); _dense(gt, "gt")
    K, N = int(num_classes), pred.shape[0]
    if out is None:
        out = torch.zeros((N, K + 1, K), dtype=torch.int64, device=pred.device)
    elif out.shape != (N, K + 1, K) or out.dtype != torch.int64 or not out.is_contiguous():
        raise ValueError("out must be a contiguous int64 [N,(K+1),K] tensor")
    _confusion_hist_batched(pred, gt, gt_lut, out, K, status_word(pred.device))
    return out
