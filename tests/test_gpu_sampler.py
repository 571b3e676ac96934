"""GPU parity: a-1 Laplace q_sample and a-2 PLMS step vs the oracle (through the C ABI)."""
import numpy as np
import pytest
import torch

from oracle import laplace as olap
from oracle.scheduler import PNDMOracle

pytestmark = pytest.mark.gpu

SHAPES = [(1, 4, 64, 64), (8, 4, 128, 128), (2, 4, 8, 8), (1, 3, 5, 7), (1, 1, 1, 1)]


def _ops():
    from ldiffusion_b200 import ops
    return ops


@pytest.mark.parametrize("shape", SHAPES)
def test_qsample_injected_noise_bit_exact(shape):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(shape, generator=g) * 5.5
    noise = torch.randn(shape, generator=g)
    want = olap.qsample_injected(x, noise)
    got = _ops().laplace_qsample(x.cuda(), 0.5, noise=noise.cuda()).cpu()
    assert torch.equal(got, want)


@pytest.mark.parametrize("shape", SHAPES[:3])
def test_qsample_injected_uniform(shape):
    """Transform on device: log1p differs from the host libm by <= 2 ulp; the
    contract is 1e-3 relative (north_star), checked at 1e-5."""
    g = torch.Generator().manual_seed(2)
    x = torch.randn(shape, generator=g) * 5.5
    u = torch.empty(shape).uniform_(torch.finfo(torch.float32).eps - 1, 1, generator=g)
    b = olap.laplace_scale(601)
    noise_want = olap.laplace_from_uniform_chain(u, b)
    got, nz = _ops().laplace_qsample(x.cuda(), b.item(), u=u.cuda(), return_noise=True)
    torch.testing.assert_close(nz.cpu(), noise_want, rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(got.cpu(), x + noise_want, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("n", [1, 3, 8, 1000, 65536 + 5])
def test_qsample_philox_stream_matches_restatement(n):
    """The kernel's own counter-based stream is reproducible on the CPU.  Philox mode uses a
    short log1p (hardware log2 / 5-term series, <= 6e-6 relative; contract 1e-3)."""
    seed, offset, b = 0x1234ABCD5678, 77, 0.75
    x = torch.zeros(n)
    got, nz = _ops().laplace_qsample(x.cuda(), b, seed=seed, offset=offset, return_noise=True)
    want = olap.laplace_philox(n, b, seed, offset)
    torch.testing.assert_close(nz.cpu(), want, rtol=2e-5, atol=1e-7)
    assert torch.equal(got, nz)                      # x == 0 -> out == noise


@pytest.fixture
def philox_rounds():
    from ldiffusion_b200 import _cabi
    lib = _cabi.lib()
    yield lambda r: lib.ldiff_tune(_cabi.TUNE_PHILOX_ROUNDS, r)
    lib.ldiff_tune(_cabi.TUNE_PHILOX_ROUNDS, 0)


@pytest.mark.parametrize("rounds", [10, 7])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_qsample_philox_distribution(philox_rounds, rounds, dtype):
    """Both round counts of the stream (10: the default, 7: LDIFF_TUNE_PHILOX_ROUNDS), both storage types (fp32: a
    24-bit uniform per word; bf16: two 16-bit variates per word with the refined top cell): E|x| = b, Var = 2 b^2, symmetric, tail bounded by b * 23 ln 2, Kolmogorov-Smirnov against the Laplace
    CDF, tail mass beyond 6 b, independence of disjoint counter ranges."""
    philox_rounds(rounds)
    b, n = 0.9166153, 1 << 22
    _, nz = _ops().laplace_qsample(torch.zeros(n, device="cuda", dtype=dtype), b, seed=7, return_noise=True)
    nz = nz.double().cpu().numpy()
    tol = 1.0 if dtype == torch.float32 else 1.5            # bf16 storage rounds every variate to 8 bits
    assert abs(np.abs(nz).mean() / b - 1) < 5e-3 * tol
    assert abs(nz.var() / (2 * b * b) - 1) < 1e-2 * tol
    assert abs(nz.mean()) < 5e-3
    assert np.abs(nz).max() <= -b * np.log(2.0 ** (-23 if dtype == torch.float32 else -38)) * (1 + 2.0 ** -7)
    xs = np.sort(nz)
    cdf = np.where(xs < 0, 0.5 * np.exp(xs / b), 1 - 0.5 * np.exp(-xs / b))
    ks = np.abs(cdf - (np.arange(n) + 0.5) / n).max()
    assert ks < (2.0 / np.sqrt(n) if dtype == torch.float32 else 6e-3)   # (bf16: the storage grid itself is a 2^-9 step)
    tail = (np.abs(nz) > 6 * b).mean()
    assert abs(tail / np.exp(-6.0) - 1) < 0.05                           # P(|x| > 6b) = e^-6
    _, nz2 = _ops().laplace_qsample(torch.zeros(n, device="cuda", dtype=dtype), b, seed=7, offset=n // 4, return_noise=True)
    assert abs(np.corrcoef(nz, nz2.double().cpu().numpy())[0, 1]) < 5e-3
    # and the stream is the restated one for this round count
    want = olap.laplace_philox(4096, b, 7, 0, rounds=rounds, storage="f32" if dtype == torch.float32 else "bf16")
    if dtype == torch.float32:
        torch.testing.assert_close(torch.from_numpy(nz[:4096]).float(), want, rtol=2e-5, atol=1e-7)
    else:
        assert ((torch.from_numpy(nz[:4096]).float() - want).abs() <= want.abs() * (2.0 ** -8 + 2e-5) + 3e-7).all()


def test_qsample_bf16_storage():
    g = torch.Generator().manual_seed(3)
    x = (torch.randn(2, 4, 64, 64, generator=g) * 5.5).bfloat16()
    noise = torch.randn(2, 4, 64, 64, generator=g).bfloat16()
    want = (x.float() + noise.float()).bfloat16()    # fp32 add, one rounding
    got = _ops().laplace_qsample(x.cuda(), 0.5, noise=noise.cuda()).cpu()
    assert torch.equal(got, want)


def test_qsample_philox_bf16_storage():
    """bf16 storage, Philox mode: the hardware log2 for every |u| (no small-|u| series; 1.7e-7*b
    absolute in the noise) and one bf16 rounding — the restated variate to a bf16 ulp."""
    n, seed, offset, b = 65536 + 24, 99, 3, 0.9166153
    got, nz = _ops().laplace_qsample(torch.zeros(n, device="cuda", dtype=torch.bfloat16), b, seed=seed,
                                     offset=offset, return_noise=True)
    want = olap.laplace_philox(n, b, seed, offset, storage="bf16")   # two 16-bit variates per word
    err = (nz.float().cpu() - want).abs()
    assert (err <= want.abs() * (2.0 ** -8 + 2e-5) + 3e-7).all()
    big = want.abs() > 1e-6
    assert torch.equal(got, nz) and (nz.float().cpu().sign()[big] == want.sign()[big]).all()


@pytest.mark.parametrize("n_set,shape", [(1, (1, 4, 64, 64)), (4, (2, 4, 64, 64)), (5, (8, 4, 128, 128)),
                                         (10, (1, 4, 16, 16)), (50, (1, 4, 8, 8)), (4, (1, 3, 5, 7))])
def test_plms_loop_bit_exact(n_set, shape):
    """Whole sampling loop (UNet replaced by seeded tensors): every latent after
    every step equals the oracle bit for bit in fp32."""
    from ldiffusion_b200 import LaplacePLMSScheduler
    g = torch.Generator().manual_seed(10 + n_set)
    ref = PNDMOracle(); ref.set_timesteps(n_set)
    sch = LaplacePLMSScheduler(); sch.set_timesteps(n_set, device="cuda")
    assert sch.timesteps.tolist() == ref.timesteps.tolist()
    x = torch.randn(shape, generator=g) * 5.5
    xd = x.cuda()
    for t_ref, t_dev in zip(ref.timesteps, sch.timesteps):
        eps = torch.randn(shape, generator=g)
        x = ref.step(eps, t_ref, ref.scale_model_input(x, t_ref))
        xd = sch.step(eps.cuda(), t_dev, sch.scale_model_input(xd, t_dev)).prev_sample
        assert torch.equal(xd.cpu(), x), f"mismatch at t={int(t_ref)}"


def test_plms_bf16_storage():
    from ldiffusion_b200 import LaplacePLMSScheduler
    g = torch.Generator().manual_seed(4)
    ref = PNDMOracle(); ref.set_timesteps(5)
    sch = LaplacePLMSScheduler(); sch.set_timesteps(5)
    shape = (2, 4, 32, 32)
    x = (torch.randn(shape, generator=g) * 5.5).bfloat16()
    xd = x.cuda()
    for t in ref.timesteps:
        eps = torch.randn(shape, generator=g).bfloat16()
        x = ref.step(eps.float(), t, x.float()).bfloat16()   # fp32 math on bf16 storage, one rounding
        xd = sch.step(eps.cuda(), int(t), xd).prev_sample
        assert torch.equal(xd.cpu(), x)
