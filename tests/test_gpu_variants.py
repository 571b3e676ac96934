"""GPU parity of the kernel forms selectable through ``ldiff_tune`` (include/ldiff.h): every decode-tail
staging shape and every lift+argmax variant must produce the same bytes; plus the training-caller kernels
(lift adjoint, warm-up step).  All of them ship in the library, so all of them run in the default GPU suite."""
import numpy as np
import pytest
import torch

from oracle import decode_tail as odt
from oracle import head as ohead

pytestmark = [pytest.mark.gpu]


@pytest.fixture
def tune():
    from ldiffusion_b200 import _cabi
    lib = _cabi.lib()
    yield lambda knob, value: lib.ldiff_tune(knob, value)
    for knob, default in ((_cabi.TUNE_ARGMAX_VARIANT, 0), (_cabi.TUNE_DECODE_TAIL_SMS, 0), (_cabi.TUNE_DECODE_TAIL_TMA, 6)):
        lib.ldiff_tune(knob, default)


@pytest.mark.timeout(120)
@pytest.mark.parametrize("shape", [(1, 3, 64, 64), (2, 3, 128, 96), (8, 3, 1024, 1024), (3, 3, 64, 192)])
@pytest.mark.parametrize("want_rgb", [False, True])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_decode_tail_tma_bit_exact(tune, shape, want_rgb, dtype):
    """Images whose planes are a multiple of 4096 pixels take the bulk-TMA staged kernel."""
    from ldiffusion_b200 import _cabi, ops
    g = torch.Generator().manual_seed(sum(shape))
    img = torch.empty(shape).uniform_(-1.3, 1.3, generator=g).to(dtype)
    tune(_cabi.TUNE_DECODE_TAIL_TMA, 0)
    rgb0, gray0 = ops.decode_tail_gray(img.cuda(), want_rgb=want_rgb)
    for variant in range(1, 14):
        tune(_cabi.TUNE_DECODE_TAIL_TMA, variant)
        rgb1, gray1 = ops.decode_tail_gray(img.cuda(), want_rgb=want_rgb)
        torch.cuda.synchronize()
        assert torch.equal(gray0, gray1), variant
        if want_rgb:
            assert torch.equal(rgb0, rgb1), variant
            if shape[-1] * shape[-2] <= 128 * 96:
                assert np.array_equal(rgb1.cpu().numpy(), odt.decode_tail_chain(img))
    # slot of a pixel-vector tensor (strided gray planes)
    planes = torch.zeros(shape[0], 3, shape[2], shape[3], dtype=torch.uint8, device="cuda")
    ops.decode_tail_gray(img.cuda(), want_rgb=False, gray_out=planes[:, 1])
    assert torch.equal(planes[:, 1], gray0) and int(planes[:, 0].max()) == 0 and int(planes[:, 2].max()) == 0


@pytest.mark.timeout(120)
@pytest.mark.parametrize("variant", [0, 1, 4, 5])
def test_lift_argmax_variants_bit_exact(tune, variant):
    """0 = envelope kernel, row form for x32 horizontal lifts (default), 1 = envelope kernel, column form always,
    4 / 5 = per-pixel evaluation with 2 / 1 columns per thread."""
    from ldiffusion_b200 import _cabi, ops
    g = torch.Generator().manual_seed(3)
    tune(_cabi.TUNE_ARGMAX_VARIANT, variant)
    for K, shape, size in ((11, (2, 32, 32), (1024, 1024)), (6, (1, 16, 16), (512, 512)), (7, (3, 8, 8), (128, 96)),
                           (15, (1, 8, 8), (64, 66)), (3, (1, 5, 7), (45, 63)),
                           # x32 horizontal lifts the row form takes: rows not a multiple of 256, a x40 vertical lift, one
                           # and two source columns (half chunks only / a lone last chunk), K = 1 and K = 15
                           (5, (1, 5, 3), (160, 96)), (4, (2, 4, 4), (160, 128)), (1, (1, 2, 1), (64, 32)),
                           (15, (1, 3, 2), (96, 64)), (11, (1, 9, 7), (300, 224))):
        logits = torch.randn((shape[0], K) + shape[1:], generator=g)
        got = torch.full((shape[0],) + size, 0xEE, dtype=torch.uint8, device="cuda")   # poisoned: a kernel that skips
        ops._lift_argmax(logits.cuda(), got)                                            # pixels cannot pass on stale data
        assert np.array_equal(got.cpu().numpy(), ohead.lift_argmax_spec(logits.numpy(), size)), (K, shape, size)


@pytest.mark.timeout(120)
@pytest.mark.parametrize("shape,size", [((2, 3, 1024, 1024), (64, 64)), ((1, 3, 50, 40), (20, 16)), ((1, 3, 16, 16), (40, 24)),
                                        ((2, 3, 64, 64), (64, 64))])
def test_lift_backward_matches_spec_and_autograd(shape, size):
    """The adjoint kernel against its numpy spec (bit-exact where footprints are disjoint) and, end to end, the
    gradient of feature_concat_autograd against torch's autograd through ldiffusion.py:240-247."""
    from oracle import bilinear as obil
    from ldiffusion_b200 import features, ops
    g = torch.Generator().manual_seed(sum(shape))
    go = torch.randn(shape[0], 1, *size, generator=g)
    got = ops.bilinear_lift_backward(go.cuda(), shape, gray=True).cpu().numpy()
    want = obil.lift_backward_spec(go.numpy(), shape, gray=True)
    disjoint = shape[2] % size[0] == 0 and shape[3] % size[1] == 0 and shape[2] // size[0] >= 2
    if disjoint or shape[2:] == size:
        assert np.array_equal(got, want)
    else:
        np.testing.assert_allclose(got, want, rtol=0, atol=5e-7 * float(np.abs(want).max()))
    rgb = torch.randn(shape[0], 3, *size, generator=g)
    np.testing.assert_allclose(ops.bilinear_lift_backward(rgb.cuda(), shape).cpu().numpy(),
                               obil.lift_backward_spec(rgb.numpy(), shape), rtol=0, atol=5e-7 * float(rgb.abs().max()) * 4)
    steps = [torch.randn(shape, generator=g) for _ in range(2)]
    dev = [s_.cuda().requires_grad_(True) for s_ in steps]
    feats = features.feature_concat_autograd(dev, size)
    up = torch.randn(feats.shape, generator=g)
    feats.backward(up.cuda())
    ref = obil.feature_concat_grad_chain(steps, up, size)
    for d, r in zip(dev, ref):
        np.testing.assert_allclose(d.grad.cpu().numpy(), r.numpy(), rtol=0, atol=5e-7 * float(r.abs().max()))


@pytest.mark.timeout(300)
def test_warmup_step_runs_on_the_product_kernels():
    """ldiffusion.py:209-255 with gradients: Laplace noising, lift + adjoint, InfoNCE forward/backward, AdamW."""
    import ldiffusion_b200 as L
    from ldiffusion_b200.standin import StandInPipeline
    model = L.LDiffusionModel("unused", "tissue")
    pipe = StandInPipeline("cuda", seed=5)
    g = torch.Generator().manual_seed(1)
    image = torch.rand(2, 3, 512, 512, generator=g)
    label = torch.zeros(2, 1, 1024, 1024, dtype=torch.uint8)
    label[:, :, :512] = 1
    label[:, :, :, 768:] = 2
    proj = torch.nn.Linear(768, 768).cuda()
    params = list(pipe.unet.parameters()) + list(proj.parameters())
    opt = torch.optim.AdamW(params, lr=1e-3)
    before = [p.detach().clone() for p in params]
    losses = [float(model.warmup_step(image, label, pipe, pipe.unet, pipe.vae, proj, opt, 2, seed=3, step_index=i))
              for i in range(3)]
    assert all(np.isfinite(v) and v > 0 for v in losses)
    assert any(not torch.equal(a, b.detach()) for a, b in zip(before, params))
    L.ops.check_status("cuda")
