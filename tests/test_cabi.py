"""CPU: the C-ABI library builds, loads, exports every symbol include/ldiff.h declares,
and rejects bad arguments with the documented codes (no kernel is launched: no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "ldiff.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ldiff_\w+)\s*\(", src)))


def test_header_and_binding_table_agree():
    from ldiffusion_b200 import _cabi
    assert _declared() == sorted(_cabi.SIGNATURES)


def test_library_loads_and_exports_every_declared_symbol(built_lib):
    from ldiffusion_b200 import _cabi
    lib = _cabi.lib()
    raw = ctypes.CDLL(built_lib)
    for name in _declared():
        assert hasattr(raw, name), name
    assert lib.ldiff_abi_version() == _cabi.ABI_VERSION
    assert lib.ldiff_strerror(0) == b"ok"
    assert b"invalid" in lib.ldiff_strerror(-1)
    assert lib.ldiff_launch_count() == 0


def test_argument_errors_are_reported_without_a_gpu(built_lib):
    from ldiffusion_b200 import _cabi
    lib = _cabi.lib()
    EINVAL, EALIGN, EUNSUP = -1, -2, -4
    assert lib.ldiff_plms_step(None, None, None, None, None, 0, 1.0, 0.0, 1.0, None, 8, 0, None) == EINVAL
    assert lib.ldiff_plms_step(16, 16, None, None, None, 7, 1.0, 0.0, 1.0, 16, 8, 0, None) == EINVAL   # bad mode
    assert lib.ldiff_plms_step(16, 16, None, None, None, 2, 1.0, 0.0, 1.0, 16, 8, 0, None) == EINVAL   # missing e1
    assert lib.ldiff_plms_step(16, 24, None, None, None, 0, 1.0, 0.0, 1.0, 16, 8, 0, None) == EALIGN
    assert lib.ldiff_plms_step(16, 16, None, None, None, 0, 1.0, 0.0, 1.0, 16, 8, 9, None) == EUNSUP  # dtype
    assert lib.ldiff_plms_step(16, 16, None, None, None, 0, 1.0, 0.0, 1.0, 16, 0, 0, None) == 0        # n == 0
    assert lib.ldiff_laplace_qsample(16, 16, 16, 16, None, 1.0, 0, 0, 8, 0, None) == EINVAL            # noise and u
    qm = lib.ldiff_laplace_qsample_map
    assert qm(16, None, 16, None, None, None, 1.0, 0, 0, 64, 16, 4, 1, 0, None) == EINVAL              # no scale map
    assert qm(16, 16, 16, None, None, None, 1.0, 0, 0, 64, 16, 4, 2, 0, None) == EINVAL                # map channels 2 of 4
    assert qm(16, 16, 16, None, None, None, 1.0, 0, 0, 60, 16, 4, 1, 0, None) == EINVAL                # n % (C*plane)
    assert qm(16, 24, 16, None, None, None, 1.0, 0, 0, 64, 16, 4, 1, 0, None) == EALIGN
    assert qm(16, 16, 16, None, None, None, 1.0, 0, 0, 64, 16, 4, 1, 2, None) == EUNSUP                # u8 storage
    assert qm(16, 16, 16, None, None, None, 1.0, 0, 0, 0, 16, 4, 1, 0, None) == 0                      # empty batch
    sr = lib.ldiff_scaled_residual
    assert sr(16, 16, 16, 16, 0.0, 64, 16, 4, 1, 0, None) == EINVAL                                    # division by 0
    assert sr(16, None, 16, 16, 1.0, 64, 16, 4, 4, 0, None) == EINVAL
    assert sr(16, 16, 16, 40, 1.0, 64, 16, 4, 4, 0, None) == EALIGN
    assert lib.ldiff_decode_tail_gray(16, None, None, 1, 4, 4, 16, 0, None) == EINVAL                  # no output
    assert lib.ldiff_decode_tail_gray(16, None, 16, 1, 4, 4, 8, 0, None) == EINVAL                     # stride < H*W
    assert lib.ldiff_bilinear_lift(16, 0, 2, 4, 4, 32, 16, 16, 0, 1, 0, 8, 8, 1, 1, None) == EINVAL    # gray needs C==3
    assert lib.ldiff_bilinear_lift(16, 2, 1, 4, 4, 16, 16, 16, 1, 1, 0, 8, 8, 1, 0, None) == EUNSUP    # u8 -> bf16
    lb = lib.ldiff_bilinear_lift_backward
    assert lb(16, 1, 0, 8, 8, 16, 2, 16, 16, 512, 256, 1, 1, None) == EINVAL                           # gray needs C == 3
    assert lb(16, 1, 1, 8, 8, 16, 3, 16, 16, 768, 256, 1, 1, None) == EINVAL                           # channel outside grad_out
    assert lb(16, 3, 0, 8, 8, 16, 3, 16, 16, 768, 100, 1, 0, None) == EINVAL                           # channel stride < h*w
    assert lb(16, 3, 0, 8, 8, 16, 3, 16, 16, 768, 256, 0, 0, None) == 0                                # empty batch
    assert lib.ldiff_tune(7, 1) == EINVAL and lib.ldiff_tune(0, -1) == EINVAL and lib.ldiff_tune(0, 0) == 0
    assert lib.ldiff_confusion_hist(16, 16, None, 16, 64, 0, 16, None) == EINVAL                       # K < 1
    assert lib.ldiff_confusion_hist(16, 16, None, 16, 64, 200, 16, None) == EUNSUP                     # K > 128
    assert lib.ldiff_lift_argmax(16, 16, 1, 300, 4, 4, 8, 8, None) == EINVAL
    assert lib.ldiff_head_logits(16, 16, None, 16, 1, 256, 64, 1024, 1, None, 0, None) == EUNSUP       # K > 32
    assert lib.ldiff_head_logits(16, 16, None, 16, 1, 256, 8, 1024, 1, None, 4, None) == EINVAL        # clear without buffer
    sn = lib.ldiff_plms_step_noise
    assert sn(16, 16, None, None, None, 0, 1.0, 0.0, 1.0, 16, None, 16, None, None, 1.0, 0, 0, 8, 0, None) == EINVAL   # no clean
    assert sn(16, 16, None, None, None, 3, 1.0, 0.0, 1.0, 16, 16, 16, None, None, 1.0, 0, 0, 8, 0, None) == EINVAL     # missing e1, e2
    assert sn(16, 16, None, None, None, 0, 1.0, 0.0, 1.0, 16, 16, 16, 16, 16, 1.0, 0, 0, 8, 0, None) == EINVAL         # noise and u
    assert sn(16, 16, None, None, None, 0, 1.0, 0.0, 1.0, 16, 16, 24, None, None, 1.0, 0, 0, 8, 0, None) == EALIGN
    assert sn(16, 16, None, None, None, 0, 1.0, 0.0, 1.0, 16, 16, 16, None, None, 1.0, 0, 0, 0, 0, None) == 0          # n == 0
    df = lib.ldiff_decode_tail_fused
    assert df(16, None, 16, 1, 32, 32, 1024, 0, None, 0, 1, 0, None, None, None, 0, None, None) == EINVAL   # nothing fused
    assert df(16, None, 16, 1, 32, 32, 1024, 0, None, 0, 1, 0, None, None, 16, 1024, None, None) == EINVAL  # plane without label
    assert df(16, None, 16, 1, 40, 32, 1280, 0, 16, 0, 1, 0, None, None, None, 0, None, None) == EUNSUP     # H % 16
    assert df(16, None, 16, 1, 32, 32, 1024, 0, 16, 0, 2, 2, None, None, None, 0, None, None) == EINVAL     # channel outside feat
    assert df(16, None, 16, 1, 32, 32, 1024, 1, 16, 2, 1, 0, None, None, None, 0, None, None) == EUNSUP     # u8 feature
    ph = lib.ldiff_lut_paint_hist
    assert ph(16, 16, 16, None, 16, 64, 1, 8, 8, 5, None, 0, 16, None) == EINVAL                        # no gt
    assert ph(16, 16, 16, 16, 16, 64, 1, 8, 8, 16, None, 0, 16, None) == EUNSUP                         # K > 15
    assert ph(16, 16, 16, 16, 16, 60, 1, 8, 8, 5, None, 0, 16, None) == EALIGN                          # n % 16
    ah = lib.ldiff_lift_argmax_hist
    assert ah(16, 16, None, 16, 1, 5, 4, 4, 64, 64, None, 0, 16, None) == EINVAL                        # no gt
    assert ah(16, 16, 16, 16, 1, 16, 4, 4, 64, 64, None, 0, 16, None) == EUNSUP                         # K > 15
    assert ah(16, 16, 16, 16, 1, 5, 4, 4, 8, 8, None, 0, 16, None) == EUNSUP                            # lift below 4x
    assert lib.ldiff_launch_count() == 0


def test_ops_refuse_cpu_tensors():
    import torch
    from ldiffusion_b200 import ops, _cabi
    with pytest.raises(_cabi.LdiffError):
        ops.plms_step(torch.zeros(8), [torch.zeros(8)], 0, 1.0, 0.1, 1.0)
    with pytest.raises(_cabi.LdiffError):
        ops.confusion_hist(torch.zeros(16, dtype=torch.uint8), torch.zeros(16, dtype=torch.uint8), 3)
    with pytest.raises(TypeError):
        ops.confusion_hist(torch.zeros(16), torch.zeros(16), 3)


def test_custom_ops_are_registered():
    import torch
    import ldiffusion_b200  # noqa: F401
    for name in ("laplace_qsample", "laplace_qsample_map", "scaled_residual", "plms_step", "plms_step_noise", "decode_tail_gray", "decode_tail_fused", "bilinear_lift", "head_logits", "lift_argmax",
                 "cell_classify", "lut_paint", "argmax_channels", "confusion_hist", "confusion_hist_batched"):
        assert hasattr(torch.ops.ldiff, name), name


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from ldiffusion_b200 import _cabi
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_cabi.LdiffError, match="no CPU fallback"):
        _cabi.lib()
