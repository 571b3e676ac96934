"""GPU: the edges of the domain — empty batches, ragged sizes that leave the vector paths, the largest class counts
each kernel takes, misaligned views — through the same public wrappers as everything else (C ABI underneath)."""
import numpy as np
import pytest
import torch

from oracle import decode_tail as odt
from oracle import head as ohead
from oracle import metrics as omet
from oracle import scheduler as osch

pytestmark = pytest.mark.gpu


def _ops():
    from ldiffusion_b200 import ops
    return ops


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_empty_batches_are_no_ops(dtype):
    ops = _ops()
    dev = "cuda"
    x = torch.zeros(0, 4, 8, 8, device=dev, dtype=dtype)
    assert ops.laplace_qsample(x, 0.5, seed=1).shape == x.shape
    assert ops.plms_step(x, [x, x], 2, 1.0, -0.1, 0.5).shape == x.shape
    img = torch.zeros(0, 3, 64, 64, device=dev, dtype=dtype)
    rgb, gray = ops.decode_tail_gray(img, want_rgb=True)
    assert rgb.shape == (0, 64, 64, 3) and gray.shape == (0, 64, 64)
    assert ops.bilinear_lift(img, (4, 4)).shape == (0, 3, 4, 4)
    C = ops.confusion_hist(torch.zeros(0, dtype=torch.uint8, device=dev), torch.zeros(0, dtype=torch.uint8, device=dev), 5)
    assert C.shape == (6, 5) and int(C.sum()) == 0
    lut = ops.cell_classify(torch.zeros(0, 256, device=dev, dtype=dtype), torch.zeros(5, 256, device=dev, dtype=dtype),
                            None, torch.zeros(0, dtype=torch.int32, device=dev), 7)
    assert lut.shape == (7,) and not lut.any()
    m = ops.lut_paint(torch.zeros(0, 16, 16, dtype=torch.int32, device=dev), torch.zeros(4, dtype=torch.uint8, device=dev))
    assert m.shape == (0, 16, 16)
    ops.check_status(torch.device(dev))


@pytest.mark.parametrize("shape", [(1, 3, 1, 1), (1, 3, 5, 7), (2, 3, 3, 16), (1, 3, 17, 31), (3, 3, 16, 16)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_decode_tail_ragged_planes(shape, dtype):
    """Planes that are not a multiple of 16 pixels leave the vector kernels (scalar path); results stay integer-exact."""
    g = torch.Generator().manual_seed(sum(shape))
    img = torch.empty(shape).uniform_(-1.4, 1.4, generator=g).to(dtype)
    rgb, gray = _ops().decode_tail_gray(img.cuda(), want_rgb=True)
    want = odt.decode_tail_chain(img)
    assert np.array_equal(rgb.cpu().numpy(), want)
    assert np.array_equal(gray.cpu().numpy(), odt.gray_chain(want))


@pytest.mark.parametrize("n", [1, 7, 9, 4097])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_sampler_kernels_ragged_lengths(n, dtype):
    """Lengths that are not a multiple of the 8-element vector: the scalar tails give the same bits as the vector body."""
    ops = _ops()
    g = torch.Generator().manual_seed(n)
    x = torch.randn(n, generator=g).to(dtype).cuda()
    e = [torch.randn(n, generator=g).to(dtype).cuda() for _ in range(4)]
    nz = torch.randn(n, generator=g).to(dtype).cuda()
    got = ops.laplace_qsample(x, 0.7, noise=nz)
    assert torch.equal(got.float(), (x.float() + nz.float()).to(dtype).float())
    big = torch.cat([x, torch.zeros(8 - n % 8, device="cuda", dtype=dtype)]) if n % 8 else x
    ebig = [torch.cat([t, torch.zeros(8 - n % 8, device="cuda", dtype=dtype)]) if n % 8 else t for t in e]
    a = ops.plms_step(x, e, 4, 0.9, -0.2, 0.7)
    b = ops.plms_step(big, ebig, 4, 0.9, -0.2, 0.7)[:n]
    assert torch.equal(a, b)
    pa = ops.laplace_qsample(x, 0.7, seed=3, offset=5)
    pb = ops.laplace_qsample(big, 0.7, seed=3, offset=5)[:n]
    assert torch.equal(pa, pb)                                   # the Philox stream is indexed by element, not by thread


def test_confusion_tiny_and_misaligned_views():
    ops = _ops()
    rng = np.random.default_rng(0)
    for n, K in ((1, 2), (3, 5), (15, 11), (17, 11), (4099, 128)):
        pred = rng.integers(0, K, n).astype(np.uint8)
        gt = rng.integers(0, K + 3, n).astype(np.uint8)
        buf_p = torch.zeros(n + 3, dtype=torch.uint8, device="cuda")
        buf_g = torch.zeros(n + 5, dtype=torch.uint8, device="cuda")
        buf_p[3:] = torch.from_numpy(pred).cuda()                # views that start off the 16-byte grid
        buf_g[5:] = torch.from_numpy(gt).cuda()
        C = ops.confusion_hist(buf_p[3:], buf_g[5:], K)
        assert np.array_equal(C.cpu().numpy(), omet.confusion_matrix(pred, gt, K))
    ops.check_status(torch.device("cuda"))


@pytest.mark.parametrize("K", [1, 2, 15, 16, 33, 255])
def test_lift_argmax_class_count_limits(K):
    """K = 1 (a constant mask) up to the uint8 limit: 15 is the envelope kernel's last K, above it the per-pixel and
    the generic kernels take over — all bit-exact against the pinned rule."""
    g = np.random.default_rng(K)
    logits = g.standard_normal((1, K, 4, 5)).astype(np.float32)
    size = (32, 40)
    got = _ops().lift_argmax(torch.from_numpy(logits).cuda(), size).cpu().numpy()
    assert np.array_equal(got, ohead.lift_argmax_spec(logits, size))


def test_lift_argmax_identity_and_down_sizes():
    """Output sizes equal to / smaller than the logits (no lift at all, a down-sample): generic kernel, exact."""
    g = np.random.default_rng(5)
    logits = g.standard_normal((2, 6, 16, 12)).astype(np.float32)
    for size in ((16, 12), (8, 6), (5, 7), (16, 48)):
        got = _ops().lift_argmax(torch.from_numpy(logits).cuda(), size).cpu().numpy()
        assert np.array_equal(got, ohead.lift_argmax_spec(logits, size)), size


def test_cell_classify_largest_k_and_feature_widths():
    ops = _ops()
    g = torch.Generator().manual_seed(2)
    for N, K, Cin, dtype in ((37, 32, 8, torch.float32), (5, 32, 1024, torch.float32), (260, 16, 512, torch.bfloat16),
                             (129, 2, 64, torch.bfloat16), (64, 17, 256, torch.bfloat16)):
        feats = torch.randn(N, Cin, generator=g).to(dtype)
        w = (torch.randn(K, Cin, generator=g) / Cin ** 0.5).to(dtype)
        b = torch.randn(K, generator=g) * 0.1
        ids = torch.arange(1, N + 1, dtype=torch.int32)
        lut, lo = ops.cell_classify(feats.cuda(), w.cuda(), b.cuda(), ids.cuda(), N + 1, return_logits=True)
        _, ref = ohead.cell_classify_chain(feats.float(), w.float(), b)
        torch.testing.assert_close(lo.cpu(), ref, rtol=1e-4, atol=1e-4)
        p = torch.softmax(lo.cpu(), 1)[:, 1:]
        assert np.array_equal(lut.cpu().numpy()[1:], (p.argmax(1) + 1).numpy().astype(np.uint8))
    ops.check_status(torch.device("cuda"))


def test_scheduler_single_step_and_long_loops():
    """n = 1 inference step (two calls, the duplicated second timestep) and a 200-step loop against the oracle."""
    from ldiffusion_b200 import LaplacePLMSScheduler
    g = torch.Generator().manual_seed(4)
    for n_set in (1, 2, 200):
        s, o = LaplacePLMSScheduler(), osch.PNDMOracle()
        s.set_timesteps(n_set, device="cuda")
        o.set_timesteps(n_set)
        x = torch.randn(1, 4, 8, 8, generator=g)
        xo, xs = x.clone(), x.cuda()
        for t, to in zip(s.timesteps, o.timesteps.tolist()):
            eps = torch.randn(1, 4, 8, 8, generator=g)
            xs = s.step(eps.cuda(), t, xs).prev_sample
            xo = o.step(eps, to, xo)
        assert torch.equal(xs.cpu(), xo), n_set
