"""GPU parity: a-3 decode tail / gray / pixel vectors and a-4 bilinear lift vs the oracle."""
import numpy as np
import pytest
import torch

from oracle import bilinear as obil
from oracle import decode_tail as odt

pytestmark = pytest.mark.gpu


def _ops():
    from ldiffusion_b200 import ops
    return ops


def _decoded(shape, seed, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    x = torch.empty(shape).uniform_(-1.2, 1.2, generator=g)        # exercises the clamp
    # sprinkle exact .5 quantisation boundaries (round-half-even) and the clamp edges
    flat = x.view(-1)
    k = torch.arange(0, min(flat.numel(), 512))
    flat[k] = ((k % 256).float() + 0.5) / 255.0 * 2 - 1
    flat[-4:] = torch.tensor([-1.0, 1.0, -5.0, 5.0])
    return x.to(dtype)


@pytest.mark.parametrize("shape", [(1, 3, 512, 512), (2, 3, 64, 48), (1, 3, 17, 13), (3, 3, 4, 4)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_decode_tail_bit_exact(shape, dtype):
    img = _decoded(shape, 5, dtype)
    want_rgb = odt.decode_tail_chain(img)
    want_gray = odt.gray_chain(want_rgb)
    assert np.array_equal(want_gray, odt.gray_spec(want_rgb))
    rgb, gray = _ops().decode_tail_gray(img.cuda())
    assert np.array_equal(rgb.cpu().numpy(), want_rgb)
    assert np.array_equal(gray.cpu().numpy(), want_gray)
    # gray only / rgb only variants
    _, gray2 = _ops().decode_tail_gray(img.cuda(), want_rgb=False)
    rgb2, _ = _ops().decode_tail_gray(img.cuda(), want_gray=False)
    assert torch.equal(gray2, gray) and torch.equal(rgb2, rgb)


@pytest.mark.parametrize("shape", [(2, 3, 64, 48), (1, 3, 17, 13)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_decode_tail_fused_model_input(shape, dtype):
    """N1: PIL -> ToTensor -> Normalize(ImageNet) hand-off fused into the decode tail, bit-exact."""
    img = _decoded(shape, 9, dtype)
    want_rgb = odt.decode_tail_chain(img)
    want = odt.model_input_chain(want_rgb)
    rgb, gray, mi = _ops().decode_tail_model_input(img.cuda(), want_gray=True)
    assert np.array_equal(rgb.cpu().numpy(), want_rgb)
    assert np.array_equal(gray.cpu().numpy(), odt.gray_spec(want_rgb))
    assert torch.equal(mi.cpu(), want)
    _, _, mi2 = _ops().decode_tail_model_input(img.cuda(), want_rgb=False, mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5))
    assert torch.equal(mi2.cpu(), odt.model_input_chain(want_rgb, (0.5, 0.5, 0.5), (0.5, 0.5, 0.5)))


def test_pixel_vectors_match_reference_loop():
    """pixel_latent_vector.py:84-93 incl. the literal per-pixel dict loop (small image)."""
    from ldiffusion_b200 import pixel_vectors
    n, H, W = 5, 32, 48
    steps = [_decoded((1, 3, H, W), 20 + i) for i in range(n)]
    label = torch.randint(0, 11, (H, W), generator=torch.Generator().manual_seed(9), dtype=torch.uint8)
    pv = pixel_vectors([s.cuda() for s in steps], label.cuda().unsqueeze(0), return_rgb=True)
    vec = pv.vectors[0].cpu().numpy()
    assert vec.shape == (H, W, n + 1)
    assert np.array_equal(vec, odt.pixel_vectors_chain(steps, label.numpy()))
    d = odt.pixel_vectors_loop(steps, label.numpy())
    for (i, j), v in list(d.items())[:: 37]:
        assert vec[i, j].tolist() == [int(a) for a in v]
    assert np.array_equal(pv.rgb.cpu().numpy(), odt.decode_tail_chain(steps[-1]))


def test_pixel_vectors_full_size_checksum():
    """cfg-2 shape [8,3,1024,1024] x 5 steps: compare against numpy on the host
    (vectorised chain, no PIL loop) through a checksum of checksums and exactly."""
    from ldiffusion_b200 import pixel_vectors
    steps = [_decoded((8, 3, 1024, 1024), 40 + i, torch.bfloat16) for i in range(2)]
    pv = pixel_vectors([s.cuda() for s in steps])
    for i, s in enumerate(steps):
        want = odt.gray_spec(odt.decode_tail_chain(s))
        got = pv.planes[:, i].cpu().numpy()
        assert int(got.astype(np.uint64).sum()) == int(want.astype(np.uint64).sum())
        assert np.array_equal(got, want)


# ------------------------------ bilinear -----------------------------------

@pytest.mark.parametrize("shape,size", [((2, 3, 64, 64), (1024, 1024)), ((1, 11, 32, 32), (1024, 1024)),
                                        ((1, 3, 37, 53), (101, 77)), ((2, 2, 16, 16), (16, 64)),
                                        ((1, 1, 8, 8), (8, 8))])
def test_lift_up_bit_exact_vs_aten(shape, size):
    """Up-sampling: bit-identical to F.interpolate on the CPU (and to the spec)."""
    x = torch.randn(shape, generator=torch.Generator().manual_seed(6))
    want = obil.lift_chain(x, size)
    assert np.array_equal(want.numpy(), obil.lift_spec(x.numpy(), size))
    got = _ops().bilinear_lift(x.cuda(), size).cpu()
    assert torch.equal(got, want)


@pytest.mark.parametrize("shape,size", [((2, 3, 1024, 1024), (64, 64)), ((1, 2, 50, 40), (20, 16)),
                                        ((1, 3, 100, 100), (64, 100))])
def test_lift_down_spec_exact_aten_close(shape, size):
    """Down-sampling: bit-identical to the spec; ATen's CPU kernel contracts its
    4-term sum differently (<= 2 ulp), compared at the 1e-3 relative contract."""
    x = torch.randn(shape, generator=torch.Generator().manual_seed(7))
    got = _ops().bilinear_lift(x.cuda(), size).cpu()
    assert np.array_equal(got.numpy(), obil.lift_spec(x.numpy(), size))
    torch.testing.assert_close(got, obil.lift_chain(x, size), rtol=1e-3, atol=1e-6)


def test_feature_concat_training_path():
    """ldiffusion.py:240-247 at the BASELINE shape: [B,3,1024,1024] x n -> [B,n,64,64]."""
    from ldiffusion_b200 import feature_concat
    steps = [torch.randn(2, 3, 1024, 1024, generator=torch.Generator().manual_seed(30 + i)) for i in range(3)]
    got = feature_concat([s.cuda() for s in steps]).cpu()
    assert np.array_equal(got.numpy(), obil.feature_concat_spec([s.numpy() for s in steps]))
    torch.testing.assert_close(got, obil.feature_concat_chain(steps), rtol=1e-3, atol=1e-6)


def test_feature_concat_as_shipped_64():
    """As shipped the decode is already 64x64 (ldiffusion.py:200,212): scale-1 copy + gray."""
    from ldiffusion_b200 import feature_concat
    steps = [torch.randn(4, 3, 64, 64, generator=torch.Generator().manual_seed(50 + i)) for i in range(2)]
    got = feature_concat([s.cuda() for s in steps]).cpu()
    assert torch.equal(got, obil.feature_concat_chain(steps))


def test_gray_up_and_bf16():
    x = torch.randn(2, 3, 64, 64, generator=torch.Generator().manual_seed(8))
    got = _ops().bilinear_lift(x.cuda(), (256, 256), gray=True).cpu()
    want = obil.gray_weighted_spec(obil.lift_spec(x.numpy(), (256, 256)))
    assert np.array_equal(got.numpy(), want)
    xb = x.bfloat16()
    gotb = _ops().bilinear_lift(xb.cuda(), (1024, 1024)).cpu()
    wantb = torch.from_numpy(obil.lift_spec(xb.float().numpy(), (1024, 1024))).bfloat16()
    assert torch.equal(gotb, wantb)                       # fp32 math, one rounding to bf16
    ref = obil.lift_chain(x, (1024, 1024))
    rel = (gotb.float() - ref).abs().max() / ref.abs().max()
    assert rel < 2 ** -7                                   # bf16 storage vs the fp32 reference


def test_label_down_bit_exact():
    """ldiffusion.py:224-226: uint8 -> float -> bilinear -> uint8."""
    from ldiffusion_b200 import label_down
    lab = torch.randint(0, 256, (3, 1, 1024, 1024), generator=torch.Generator().manual_seed(11), dtype=torch.uint8)
    got = label_down(lab.cuda()).cpu()
    assert torch.equal(got, obil.label_down_chain(lab))
    lab2 = torch.randint(0, 256, (1, 1, 100, 60), generator=torch.Generator().manual_seed(12), dtype=torch.uint8)
    got2 = label_down(lab2.cuda(), (33, 20)).cpu()
    assert np.array_equal(got2.numpy(), obil.label_down_spec(lab2.numpy(), (33, 20)))


def test_rgb_up_full_size():
    from ldiffusion_b200 import rgb_up
    x = torch.randn(8, 3, 64, 64, generator=torch.Generator().manual_seed(13))
    got = rgb_up(x.cuda()).cpu()
    assert torch.equal(got, obil.lift_chain(x, (1024, 1024)))


def test_lift_multi_source_equals_single_launches():
    steps = [torch.randn(2, 3, 256, 256, generator=torch.Generator().manual_seed(70 + i)).cuda() for i in range(5)]
    multi = _ops().bilinear_lift_multi(steps, (64, 64), gray=True)
    single = torch.cat([_ops().bilinear_lift(s, (64, 64), gray=True) for s in steps], dim=1)
    assert torch.equal(multi, single)
    multi3 = _ops().bilinear_lift_multi(steps[:2], (32, 48))
    assert torch.equal(multi3, torch.cat([_ops().bilinear_lift(s, (32, 48)) for s in steps[:2]], dim=1))
