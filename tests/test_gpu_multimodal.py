"""GPU parity: the multimodal variant of a-1 — Laplace(0,1) noise modulated by a depth map and
its inverse (segmentor.py:339-345, :375-379) — through the C ABI, and the Segmentor method built on it."""
import numpy as np
import pytest
import torch

from oracle import laplace as olap

pytestmark = pytest.mark.gpu

# (shape, channels of the scale map): latent shape of the reference, the vector path with a full and a
# broadcast map, planes that are not a multiple of the vector width (scalar path), one element,
# vector path with planes / channel counts that are not powers of two (division instead of shifts)
CASES = [((1, 4, 32, 32), 1), ((3, 4, 32, 32), 1), ((3, 4, 32, 32), 4), ((8, 4, 128, 128), 1),
         ((2, 4, 5, 7), 1), ((2, 3, 5, 7), 3), ((1, 1, 1, 1), 1), ((2, 4, 3, 4), 1), ((2, 3, 6, 4), 1),
         ((2, 4, 6, 12), 1)]


def _ops():
    from ldiffusion_b200 import ops
    return ops


def _inputs(shape, cs, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(shape, generator=g) * 5.5
    s = torch.rand((shape[0], cs) + tuple(shape[2:]), generator=g) * 2
    return g, x, s


@pytest.mark.parametrize("shape,cs", CASES)
@pytest.mark.parametrize("x_mul", [None, 0.18215])
def test_depth_noising_injected_noise_bit_exact(shape, cs, x_mul):
    g, x, s = _inputs(shape, cs, 21)
    noise = torch.distributions.Laplace(0.0, 1.0).sample(shape)
    want, _ = olap.depth_noising_chain(x, s, noise=noise, x_mul=x_mul)
    got = _ops().laplace_qsample_map(x.cuda(), s.cuda(), noise=noise.cuda(), x_mul=1.0 if x_mul is None else x_mul)
    assert torch.equal(got.cpu(), want)


@pytest.mark.parametrize("shape,cs", CASES)
@pytest.mark.parametrize("out_div", [None, 0.18215])
def test_depth_unnoise_bit_exact(shape, cs, out_div):
    g, x, s = _inputs(shape, cs, 22)
    eps = torch.randn(shape, generator=g)
    want = olap.depth_unnoise_chain(x, eps, s, out_div=out_div)
    got = _ops().scaled_residual(x.cuda(), eps.cuda(), s.cuda(), out_div=1.0 if out_div is None else out_div)
    assert torch.equal(got.cpu(), want)


def test_depth_round_trip_is_the_identity_for_a_perfect_prediction():
    """noise_pred == noise: x -> noising -> un-noising returns x to within one rounding of x + n*s."""
    g, x, s = _inputs((4, 4, 32, 32), 1, 23)
    noise = torch.distributions.Laplace(0.0, 1.0).sample(x.shape)
    o = _ops()
    noisy = o.laplace_qsample_map(x.cuda(), s.cuda(), noise=noise.cuda())
    back = o.scaled_residual(noisy, noise.cuda(), s.cuda()).cpu()
    ulp = 1.01 * torch.finfo(torch.float32).eps * (x.abs() + (noise * s).abs())
    assert ((back - x).abs() <= ulp).all()


@pytest.mark.parametrize("shape,cs", [((2, 4, 32, 32), 1), ((2, 4, 5, 7), 1), ((1, 4, 8, 8), 4)])
def test_depth_noising_injected_uniform(shape, cs):
    """rsample's transform on the device (log1pf within 2 ulp of the host libm; contract 1e-3)."""
    g, x, s = _inputs(shape, cs, 24)
    u = torch.empty(shape).uniform_(torch.finfo(torch.float32).eps - 1, 1, generator=g)
    noise_want = olap.laplace_from_uniform_chain(u, 1.0)
    got, nz = _ops().laplace_qsample_map(x.cuda(), s.cuda(), u=u.cuda(), return_noise=True)
    torch.testing.assert_close(nz.cpu(), noise_want, rtol=1e-5, atol=1e-7)
    assert torch.equal(got.cpu(), olap.depth_noising_chain(x, s, noise=nz.cpu())[0])   # the add/mul part is exact


@pytest.mark.parametrize("shape,cs", [((2, 4, 32, 32), 1), ((1, 4, 5, 7), 1), ((1, 3, 3, 3), 3)])
def test_depth_noising_philox_stream_matches_restatement(shape, cs):
    """Same counter-based stream as ldiff_laplace_qsample (element i = word i%4 of counter offset + i//4)."""
    seed, offset = 0xBEEF1234567, 5
    g, x, s = _inputs(shape, cs, 25)
    got, nz = _ops().laplace_qsample_map(x.cuda(), s.cuda(), seed=seed, offset=offset, return_noise=True)
    want_nz = olap.laplace_philox(x.numel(), 1.0, seed, offset).reshape(shape)
    torch.testing.assert_close(nz.cpu(), want_nz, rtol=2e-5, atol=1e-7)
    assert torch.equal(got.cpu(), olap.depth_noising_chain(x, s, noise=nz.cpu())[0])
    plain = _ops().laplace_qsample(torch.zeros(x.numel(), device="cuda"), 1.0, seed=seed, offset=offset)
    assert torch.equal(plain.cpu().reshape(shape), nz.cpu())       # identical draws to the scalar-b kernel


def test_depth_ops_bf16_storage():
    """fp32 arithmetic on the bf16 values, one final rounding."""
    g, x, s = _inputs((2, 4, 32, 32), 1, 26)
    x, s = x.bfloat16(), s.bfloat16()
    noise = torch.randn(x.shape, generator=g).bfloat16()
    o = _ops()
    want = olap.depth_noising_chain(x.float(), s.float(), noise=noise.float(), x_mul=0.18215)[0].bfloat16()
    assert torch.equal(o.laplace_qsample_map(x.cuda(), s.cuda(), noise=noise.cuda(), x_mul=0.18215).cpu(), want)
    want = olap.depth_unnoise_chain(x.float(), noise.float(), s.float(), out_div=0.18215).bfloat16()
    assert torch.equal(o.scaled_residual(x.cuda(), noise.cuda(), s.cuda(), out_div=0.18215).cpu(), want)


def test_depth_ops_argument_errors():
    o = _ops()
    x = torch.zeros(2, 4, 8, 8, device="cuda")
    with pytest.raises(ValueError):
        o.laplace_qsample_map(x, torch.zeros(2, 2, 8, 8, device="cuda"))            # 2 is neither 1 nor C
    with pytest.raises(ValueError):
        o.laplace_qsample_map(x, torch.zeros(1, 1, 8, 8, device="cuda"))            # batch mismatch
    with pytest.raises(ValueError):
        o.laplace_qsample_map(x, torch.zeros(2, 1, 8, 8, device="cuda"), noise=x, u=x)
    with pytest.raises(ValueError):
        o.scaled_residual(x, x[:1], torch.zeros(2, 1, 8, 8, device="cuda"))
    with pytest.raises(ValueError):
        o.scaled_residual(x, x, torch.zeros(2, 1, 8, 8, device="cuda"), out_div=0.0)
    assert o.laplace_qsample_map(x[:0], torch.zeros(0, 1, 8, 8, device="cuda")).shape == (0, 4, 8, 8)   # empty batch


def test_multimodal_augment_matches_oracle_on_captured_tensors():
    """Segmentor.ldiffusion_augment_for_multimodal with stand-in VAE / UNet / ControlNet: capture the
    tensors either side of the two kernels and redo segmentor.py:339-345 / :375-379 on the CPU oracle."""
    import ldiffusion_b200 as L
    from ldiffusion_b200.standin import StandInControlNet, StandInPipeline
    seg = L.Segmentor(None, None, "tissue", 7)
    pipe = StandInPipeline("cuda", seed=6)
    controlnet = StandInControlNet().cuda().eval()
    cap = {}
    vae_enc, vae_dec, unet_fwd = pipe.vae.encode, pipe.vae.decode, pipe.unet.forward

    def encode(x):
        out = vae_enc(x); cap["rgb256"] = x.detach().cpu(); cap["lat"] = out.latent_dist.sample().detach().cpu(); return out

    def decode(z):
        cap["z"] = z.detach().cpu(); return vae_dec(z)

    def unet_hook(s, t, **kw):
        out = unet_fwd(s, t, **kw)
        cap["noisy"] = s.detach().cpu(); cap["pred"] = out.sample.detach().cpu(); cap["mid"] = kw["mid_block_additional_residual"]
        return out

    pipe.vae.encode, pipe.vae.decode, pipe.unet.forward = encode, decode, unet_hook
    unet = pipe.unet
    g = torch.Generator().manual_seed(8)
    rgb, dtm = torch.rand(3, 3, 300, 280, generator=g), torch.rand(3, 1, 300, 280, generator=g) * 2
    noise = torch.distributions.Laplace(0.0, 1.0).sample((3, 4, 32, 32))
    out = seg.ldiffusion_augment_for_multimodal(rgb, dtm, pipe, unet, pipe.vae, controlnet, 3, "cuda",
                                                noise=noise.cuda())
    assert len(out) == 3 and all(o.shape == (256, 256, 3) and o.dtype == np.float32 for o in out)
    assert cap["mid"] is not None and cap["rgb256"].shape == (3, 3, 256, 256)
    F = torch.nn.functional
    dtm256 = F.interpolate(dtm, size=(256, 256), mode="bilinear", align_corners=False)
    depth = F.interpolate(dtm256, size=(32, 32), mode="bilinear", align_corners=False)
    torch.testing.assert_close(cap["rgb256"], F.interpolate(rgb, size=(256, 256), mode="bilinear", align_corners=False),
                               rtol=1e-6, atol=1e-6)
    want_noisy, _ = olap.depth_noising_chain(cap["lat"], depth, noise=noise, x_mul=0.18215)
    torch.testing.assert_close(cap["noisy"], want_noisy, rtol=1e-5, atol=1e-5)     # depth: two lifts, <= 2 ulp each
    # given the captured tensors, the inverse is bit-exact
    depth_dev = L.ops.bilinear_lift(L.ops.bilinear_lift(dtm.cuda(), (256, 256)), (32, 32)).cpu()
    assert torch.equal(cap["z"], olap.depth_unnoise_chain(cap["noisy"], cap["pred"], depth_dev, out_div=0.18215))
    assert torch.equal(cap["noisy"], olap.depth_noising_chain(cap["lat"], depth_dev, noise=noise, x_mul=0.18215)[0])
    # Philox path: runs, deterministic per seed, different across seeds
    a = seg.ldiffusion_augment_for_multimodal(rgb, dtm, pipe, unet, pipe.vae, controlnet, 3, "cuda", seed=1)
    b = seg.ldiffusion_augment_for_multimodal(rgb, dtm, pipe, unet, pipe.vae, controlnet, 3, "cuda", seed=1)
    c = seg.ldiffusion_augment_for_multimodal(rgb, dtm, pipe, unet, pipe.vae, controlnet, 3, "cuda", seed=2)
    assert np.array_equal(a[0], b[0]) and not np.array_equal(a[0], c[0])
