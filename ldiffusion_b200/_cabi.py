"""ctypes binding of libldiff_sm100.so (the C ABI declared in include/ldiff.h).

There is no fallback: if the library has not been built, importing a kernel
entry raises.  Build with ``python -m ldiffusion_b200.build`` (or
``__graft_entry__.build()``).
"""
import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libldiff_sm100.so")
ABI_VERSION = 2

F32, BF16, U8 = 0, 1, 2
TUNE_ARGMAX_VARIANT, TUNE_DECODE_TAIL_SMS, TUNE_DECODE_TAIL_TMA, TUNE_PHILOX_ROUNDS = 0, 1, 2, 3
STATUS_PRED_RANGE, STATUS_INST_RANGE, STATUS_SW_INF, STATUS_XCHG_TIMEOUT, STATUS_LABEL_RANGE = 1, 2, 4, 8, 16

# name -> (restype, argtypes); mirrors include/ldiff.h one to one
SIGNATURES = {
    "ldiff_abi_version": (c_int, []),
    "ldiff_strerror": (c_char_p, [c_int]),
    "ldiff_launch_count": (c_int64, []),
    "ldiff_tune": (c_int, [c_int, c_int]),
    "ldiff_tune_get": (c_int, [c_int]),
    "ldiff_laplace_qsample": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float,
                                      c_uint64, c_uint64, c_int64, c_int, c_void_p]),
    "ldiff_laplace_qsample_map": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float,
                                          c_uint64, c_uint64, c_int64, c_int64, c_int, c_int, c_int, c_void_p]),
    "ldiff_scaled_residual": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_int64, c_int64, c_int,
                                      c_int, c_int, c_void_p]),
    "ldiff_plms_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_float,
                                c_float, c_float, c_void_p, c_int64, c_int, c_void_p]),
    "ldiff_plms_step_noise": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_float,
                                      c_float, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_float, c_uint64, c_uint64, c_int64, c_int, c_void_p]),
    "ldiff_decode_tail_gray": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int64,
                                       c_int, c_void_p]),
    "ldiff_decode_tail_model_input": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                              c_int, c_int, c_int64, c_int, c_void_p]),
    "ldiff_decode_tail_fused": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int64, c_int, c_void_p,
                                        c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int64, c_void_p,
                                        c_void_p]),
    "ldiff_bilinear_lift": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int64, c_int64, c_void_p,
                                    c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "ldiff_bilinear_lift_multi": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int64, c_int64, c_void_p,
                                          c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "ldiff_bilinear_lift_backward": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int,
                                             c_int64, c_int64, c_int, c_int, c_void_p]),
    "ldiff_head_logits": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                  c_int, c_void_p, c_int, c_void_p]),
    "ldiff_lift_argmax_hist": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                       c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "ldiff_lut_paint_hist": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int,
                                     c_int64, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "ldiff_lut_paint_hist_u16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int,
                                         c_int64, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "ldiff_lift_argmax": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                  c_void_p]),
    "ldiff_cell_classify": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_void_p,
                                    c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "ldiff_copy_planes_u8": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int64, c_void_p]),
    "ldiff_lut_paint": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int64,
                                c_void_p, c_void_p]),
    "ldiff_lut_paint_u16": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int64,
                                    c_void_p, c_void_p]),
    "ldiff_argmax_channels": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int64, c_int, c_void_p]),
    "ldiff_confusion_hist": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int,
                                     c_void_p, c_void_p]),
    "ldiff_confusion_hist_batched": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int,
                                             c_void_p, c_void_p]),
    "ldiff_xchg_create": (c_int, [c_int, c_int, c_int, c_int, c_void_p]),
    "ldiff_xchg_ipc_handle": (c_int, [c_void_p, c_void_p]),
    "ldiff_xchg_connect_ipc": (c_int, [c_void_p, c_void_p]),
    "ldiff_xchg_connect_local": (c_int, [c_void_p, c_void_p]),
    "ldiff_xchg_set_timeout": (c_int, [c_void_p, c_int64]),
    "ldiff_xchg_destroy": (c_int, [c_void_p]),
    "ldiff_confusion_hist_push": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_int,
                                          c_void_p, c_void_p]),
    "ldiff_xchg_push": (c_int, [c_void_p, c_void_p, c_void_p]),
    "ldiff_xchg_reduce": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "ldiff_labels_to_u8": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "ldiff_infonce_sample": (c_int, [c_void_p, c_int, c_int64, c_int, c_int, c_uint64, c_uint64, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ldiff_infonce_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                      c_int64, c_int, c_int, c_float, c_void_p]),
    "ldiff_infonce_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_int, c_int64, c_int, c_int, c_float, c_void_p]),
    "ldiff_sw_accumulate": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                    c_int, c_int, c_void_p]),
    "ldiff_sw_tta_merge": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p]),
    "ldiff_sw_finalize_argmax": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_void_p,
                                         c_void_p]),
}

_lib = None


class LdiffError(RuntimeError):
    pass


def lib():
    """The loaded library (loaded once).  Raises if it is missing or stale."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LdiffError(
                f"{LIB_PATH} not found: the CUDA extension is not built and there is no CPU "
                "fallback.  Run `python -m ldiffusion_b200.build`.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)      # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        if handle.ldiff_abi_version() != ABI_VERSION:
            raise LdiffError("libldiff_sm100.so ABI version mismatch: rebuild the extension")
        _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        raise LdiffError(f"ldiff: {lib().ldiff_strerror(rc).decode()} (code {rc})")


def launch_count():
    return int(lib().ldiff_launch_count())
