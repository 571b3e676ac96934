"""GPU parity of the round-2 fused kernels: each one must produce exactly the bytes of the separate
launches it replaces (which are themselves pinned against the oracle elsewhere) AND match the oracle
directly on seeded inputs.

* step_then_noise        = plms_step + laplace_qsample                 (segmentor.py:100-104 + ldiffusion.py:233-237)
* decode_tail_fused      = decode_tail_gray + bilinear_lift(gray) + copy_planes_u8 + label down-lift
                                                                        (pixel_latent_vector.py:80-93, ldiffusion.py:224-226,240-247)
* lut_paint_hist         = lut_paint + confusion_hist                  (conductor.py:224-231 + utils.py:55-104)
* lift_argmax_hist       = lift_argmax + confusion_hist                (conductor.py:135, segmentor.py:536 + utils.py:55-104)
"""
import numpy as np
import pytest
import torch

from oracle import bilinear as obil
from oracle import head as ohead
from oracle import metrics as omet
from oracle.scheduler import PNDMOracle

pytestmark = pytest.mark.gpu


def _ops():
    from ldiffusion_b200 import ops
    return ops


@pytest.fixture
def tune():
    from ldiffusion_b200 import _cabi
    lib = _cabi.lib()
    yield lambda knob, value: lib.ldiff_tune(knob, value)
    for knob, default in ((_cabi.TUNE_ARGMAX_VARIANT, 0), (_cabi.TUNE_DECODE_TAIL_SMS, 0), (_cabi.TUNE_DECODE_TAIL_TMA, 6)):
        lib.ldiff_tune(knob, default)


# ---------------------------------------------------------------- step_then_noise

@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("n", [4 * 64 * 64, 1003])
def test_step_then_noise_equals_two_launches(dtype, n):
    """Every PLMS mode x every noise source: the fused launch == the two separate ones, bit for bit."""
    ops = _ops()
    g = torch.Generator().manual_seed(n)
    x, clean = [(torch.randn(n, generator=g) * 5.5).to(dtype).cuda() for _ in range(2)]
    eps = [torch.randn(n, generator=g).to(dtype).cuda() for _ in range(4)]
    inj = torch.randn(n, generator=g).to(dtype).cuda()
    u = (torch.rand(n, generator=g) * 1.98 - 0.99).to(dtype).cuda()
    for mode in range(5):
        for src in ("philox", "noise", "u"):
            kw = {"noise": inj} if src == "noise" else ({"u": u} if src == "u" else {"seed": 77, "offset": 12345})
            prev_a = ops.plms_step(x, eps, mode, 1.01, -0.02, 0.3)
            noisy_a = ops.laplace_qsample(clean, 0.7, **kw)
            prev_b, noisy_b = ops.plms_step_noise(x, eps, mode, 1.01, -0.02, 0.3, clean, 0.7, **kw)
            assert torch.equal(prev_a, prev_b), (mode, src)
            assert torch.equal(noisy_a, noisy_b), (mode, src)


def test_scheduler_step_then_noise_loop_matches_oracle():
    """N=5 loop through LaplacePLMSScheduler.step_then_noise: latents bit-exact vs the PNDM oracle, noisy
    tensors equal the stand-alone operator's, and the device-scalar timesteps are honoured by value."""
    from ldiffusion_b200 import LaplacePLMSScheduler
    g = torch.Generator().manual_seed(9)
    shape = (2, 4, 32, 32)
    x0 = torch.randn(shape, generator=g) * 5.5
    ref = PNDMOracle(); ref.set_timesteps(4)
    eps = [torch.randn(shape, generator=g) for _ in ref.timesteps]
    x = x0
    for e, t in zip(eps, ref.timesteps):
        x = ref.step(e, t, x)
    sch = LaplacePLMSScheduler()
    sch.set_timesteps(4, device="cuda")
    xd, clean = x0.cuda(), x0.cuda()
    blocks = (x0.numel() + 3) // 4
    for i, t in enumerate(sch.timesteps):                  # CUDA 0-dim tensors, resolved without a sync
        xd, noisy = sch.step_then_noise(eps[i].cuda(), t, xd, clean, seed=3, offset=i * blocks)
        want = sch.add_laplace_noise(clean, sch._host_timesteps[i], seed=3, offset=i * blocks)
        assert torch.equal(noisy, want)
    assert torch.equal(xd.cpu(), x)


# ---------------------------------------------------------------- decode_tail_fused

@pytest.mark.parametrize("shape", [(2, 3, 64, 64), (1, 3, 48, 80), (2, 3, 1024, 1024), (1, 3, 128, 32)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("tma", [0, 1, 4, 6, 7, 9, 11, 12])
def test_decode_tail_fused_equals_separate_launches(tune, shape, dtype, tma):
    from ldiffusion_b200 import _cabi
    ops = _ops()
    tune(_cabi.TUNE_DECODE_TAIL_TMA, tma)
    B, _, H, W = shape
    fh, fw = H // 16, W // 16
    g = torch.Generator().manual_seed(H + W)
    img = torch.empty(shape).uniform_(-1.3, 1.3, generator=g).to(dtype).cuda()
    label = torch.randint(0, 256, (B, H, W), generator=g, dtype=torch.uint8).cuda()
    planes = torch.zeros(B, 3, H, W, dtype=torch.uint8, device="cuda")
    rgb = torch.zeros(B, H, W, 3, dtype=torch.uint8, device="cuda")
    feat = torch.zeros(B, 2, fh, fw, dtype=dtype, device="cuda")
    feat32 = torch.zeros(B, 2, fh, fw, dtype=torch.float32, device="cuda")
    small = torch.zeros(B, 3, fh, fw, dtype=dtype, device="cuda")
    lsmall = torch.zeros(B, 1, fh, fw, dtype=torch.uint8, device="cuda")
    ops.decode_tail_fused(img, planes[:, 0], rgb_out=rgb, feat_out=feat, feat_channel=1, small_rgb_out=small,
                          label=label, label_plane_out=planes[:, 2], label_small_out=lsmall)
    ops.decode_tail_fused(img, planes[:, 1], feat_out=feat32, feat_channel=0)       # gray only + fp32 feature
    rgb_ref, gray_ref = ops.decode_tail_gray(img, want_rgb=True)
    assert torch.equal(rgb, rgb_ref)
    assert torch.equal(planes[:, 0], gray_ref) and torch.equal(planes[:, 1], gray_ref)
    assert torch.equal(planes[:, 2], label)
    assert torch.equal(feat[:, 1:2], ops.bilinear_lift(img, (fh, fw), gray=True)) and int(feat[:, 0].abs().max()) == 0
    assert torch.equal(feat32[:, 0:1], ops.bilinear_lift(img, (fh, fw), gray=True, out_dtype=torch.float32))
    assert torch.equal(small, ops.bilinear_lift(img, (fh, fw)))
    assert torch.equal(lsmall, ops.bilinear_lift(label.unsqueeze(1), (fh, fw)))
    if dtype == torch.float32 and H * W <= 128 * 128:     # and the oracle directly
        want = obil.feature_concat_spec([img.cpu().numpy()], (fh, fw))
        assert np.array_equal(feat32[:, 0:1].cpu().numpy(), want)
        assert torch.equal(lsmall.cpu(), obil.label_down_chain(label.cpu().unsqueeze(1), (fh, fw)))


def test_decode_tail_fused_argument_errors():
    ops = _ops()
    img = torch.zeros(1, 3, 40, 64, device="cuda")
    gray = torch.zeros(1, 40, 64, dtype=torch.uint8, device="cuda")
    with pytest.raises(ValueError):
        ops.decode_tail_fused(img, gray, feat_out=torch.zeros(1, 1, 2, 4, device="cuda"))    # H % 16 != 0
    img = torch.zeros(1, 3, 32, 64, device="cuda")
    gray = torch.zeros(1, 32, 64, dtype=torch.uint8, device="cuda")
    with pytest.raises(ValueError):
        ops.decode_tail_fused(img, gray, feat_out=torch.zeros(1, 1, 4, 4, device="cuda"))    # wrong feature size
    with pytest.raises(ValueError):
        ops.decode_tail_fused(img, gray, label_plane_out=gray.clone())                       # no label


# ---------------------------------------------------------------- lut_paint_hist

@pytest.mark.parametrize("K,B,H,W", [(11, 2, 256, 256), (6, 1, 64, 48), (15, 3, 128, 128), (20, 1, 64, 64), (11, 1, 30, 30)])
def test_lut_paint_hist_equals_separate_launches(K, B, H, W):
    ops = _ops()
    rng = np.random.default_rng(K * H)
    N = 200
    inst = torch.from_numpy(rng.integers(0, N + 1, (B, H, W)).astype(np.int32)).cuda()
    lut = torch.from_numpy(rng.integers(0, K, (B, N + 1)).astype(np.uint8)).cuda()
    gt = rng.integers(0, K + 2, (B, H, W)).astype(np.uint8)
    gt[gt >= K] = 255
    gtd = torch.from_numpy(gt).cuda()
    C = torch.zeros(K + 1, K, dtype=torch.int64, device="cuda")
    mask, C = ops.lut_paint_hist(inst, lut, gtd, K, out=C)
    want_mask = ops.lut_paint(inst, lut)
    assert torch.equal(mask, want_mask)
    assert np.array_equal(C.cpu().numpy(), omet.confusion_matrix(want_mask.cpu().numpy(), gt, K))
    mask2, C = ops.lut_paint_hist(inst, lut, gtd, K, out=C)                                    # accumulates
    assert np.array_equal(C.cpu().numpy(), 2 * omet.confusion_matrix(want_mask.cpu().numpy(), gt, K))
    ops.check_status("cuda")


@pytest.mark.parametrize("K,B,H,W", [(11, 2, 256, 256), (6, 1, 64, 48), (20, 1, 64, 64), (11, 1, 30, 30)])
def test_lut_paint_uint16_instance_maps(K, B, H, W):
    """Cellpose's label image is uint16 below 65 536 labels (conductor.py:180): painting and the fused histogram
    take it as it is and give what the int32 form and the oracle give — also for ids above 32 767 (an int16 view
    of the same bytes is read as unsigned) and for ids outside the LUT."""
    ops = _ops()
    rng = np.random.default_rng(K * H + 1)
    N = 40000
    ids = rng.integers(0, N + 1, (B, H, W))
    ids[:, 0, :8] = [0, 1, 32767, 32768, 39999, N, 5, 0]
    lut_np = rng.integers(0, K, (B, N + 1)).astype(np.uint8)
    lut = torch.from_numpy(lut_np).cuda()
    gt = rng.integers(0, K + 2, (B, H, W)).astype(np.uint8)
    gt[gt >= K] = 255
    gtd = torch.from_numpy(gt).cuda()
    i32 = torch.from_numpy(ids.astype(np.int32)).cuda()
    u16 = torch.from_numpy(ids.astype(np.uint16)).cuda()
    want = np.stack([ohead.cell_paint_spec(ids[b], lut_np[b]) for b in range(B)])
    for inst in (u16, u16.view(torch.int16)):
        assert np.array_equal(ops.lut_paint(inst, lut).cpu().numpy(), want)
        mask, C = ops.lut_paint_hist(inst, lut, gtd, K)
        assert np.array_equal(mask.cpu().numpy(), want)
        assert np.array_equal(C.cpu().numpy(), omet.confusion_matrix(want, gt, K))
    assert torch.equal(ops.lut_paint(i32, lut), ops.lut_paint(u16, lut))
    ops.check_status("cuda")
    ops.lut_paint(u16, lut[:, :1000].contiguous())             # ids outside a shorter LUT: class 0 + a status bit
    with pytest.raises(RuntimeError):
        ops.check_status("cuda")
    with pytest.raises(TypeError):
        ops.lut_paint(i32.long(), lut)


def test_lut_paint_hist_range_errors():
    ops = _ops()
    inst = torch.tensor([[0, 1, 5, 2]], dtype=torch.int32).repeat(4, 4).cuda()
    lut = torch.tensor([0, 3, 4], dtype=torch.uint8).cuda()
    gt = torch.zeros(4, 16, dtype=torch.uint8, device="cuda")
    ops.lut_paint_hist(inst.view(1, 4, 16), lut, gt.view(1, 4, 16), 5)
    with pytest.raises(RuntimeError):                      # instance id 5 outside the LUT
        ops.check_status("cuda")
    ops.lut_paint_hist(inst.clamp(max=2).view(1, 4, 16), lut, gt.view(1, 4, 16), 4)   # lut holds class 4 >= K
    with pytest.raises(RuntimeError):
        ops.check_status("cuda")


# ---------------------------------------------------------------- lift_argmax (envelope kernel) + hist

def _smooth_logits(B, K, h, w, seed):
    """Spatially smooth maps (what a trained decoder emits): low-frequency waves per class."""
    rng = np.random.default_rng(seed)
    yy, xx = np.meshgrid(np.linspace(0, 1, h), np.linspace(0, 1, w), indexing="ij")
    out = np.zeros((B, K, h, w), np.float32)
    for b in range(B):
        for k in range(K):
            fy, fx, ph = rng.uniform(0.5, 2.5, 2).tolist() + [rng.uniform(0, 6.28)]
            out[b, k] = 3 * np.sin(2 * np.pi * (fy * yy + fx * xx) + ph) + rng.normal(0, 0.05, (h, w))
    return out


CASES = {
    "iid": lambda: np.random.default_rng(1).standard_normal((2, 11, 32, 32)).astype(np.float32),
    "smooth": lambda: _smooth_logits(2, 11, 32, 32, 2),
    # adversarial: every class within 1e-5 of the others everywhere -> all pixels are near-ties (queue overflow path)
    "within_1e-5": lambda: (1.5 + np.random.default_rng(3).uniform(-5e-6, 5e-6, (1, 11, 8, 8))).astype(np.float32),
    # near-parallel lines: two classes differ by ulps at both corners of every column
    "ulp_pairs": lambda: (lambda a: np.concatenate([a, np.nextafter(a, np.float32(9)), a - 1], 1))(
        np.random.default_rng(4).standard_normal((1, 1, 8, 8)).astype(np.float32) * 3),
    "constant": lambda: np.zeros((1, 6, 4, 4), np.float32),
    "big_magnitude": lambda: (np.random.default_rng(5).standard_normal((1, 7, 8, 8)) * 3e4).astype(np.float32),
    "k1": lambda: np.random.default_rng(6).standard_normal((1, 1, 4, 4)).astype(np.float32),
}


@pytest.mark.parametrize("case", sorted(CASES))
@pytest.mark.parametrize("factor,variant", [(32, 0), (32, 1), (8, 0)])
def test_envelope_lift_argmax_bit_exact_vs_spec_and_chain(tune, case, factor, variant):
    """factor 32, variant 0: the row form (lift_argmax_row.cu); variant 1 / factor 8: the column form."""
    from ldiffusion_b200 import _cabi
    ops = _ops()
    tune(_cabi.TUNE_ARGMAX_VARIANT, variant)
    logits = CASES[case]()
    B, K, h, w = logits.shape
    size = (h * factor, w * factor)
    want = ohead.lift_argmax_spec(logits, size)
    got = ops.lift_argmax(torch.from_numpy(logits).cuda(), size).cpu().numpy()
    assert np.array_equal(got, want)
    if case in ("iid", "smooth", "big_magnitude", "k1"):   # no near-ties: the literal torch chain agrees as well
        chain = ohead.lift_argmax_chain(torch.from_numpy(logits), size).numpy().astype(np.uint8)
        assert np.array_equal(got, chain)


@pytest.mark.parametrize("h,H", [(10, 1024), (8, 1000), (3, 381), (12, 1024), (7, 77)])
def test_lift_argmax_tall_bands(h, H):
    """Lift factors around 100: the first band holds ~1.5*H/h rows (the old guard let 154-row bands
    into 128-entry shared arrays).  Bands over 127 rows must take the generic kernel, not corrupt memory."""
    ops = _ops()
    logits = np.random.default_rng(h * H).standard_normal((1, 5, h, h)).astype(np.float32)
    got = ops.lift_argmax(torch.from_numpy(logits).cuda(), (H, 64)).cpu().numpy()
    assert np.array_equal(got, ohead.lift_argmax_spec(logits, (H, 64)))


@pytest.mark.parametrize("case", sorted(CASES))
@pytest.mark.parametrize("variant", [0])
def test_lift_argmax_hist_equals_separate_launches(tune, case, variant):
    from ldiffusion_b200 import _cabi
    ops = _ops()
    tune(_cabi.TUNE_ARGMAX_VARIANT, variant)
    logits = CASES[case]()
    B, K, h, w = logits.shape
    size = (h * 32, w * 32)
    rng = np.random.default_rng(7)
    gt = rng.integers(0, K + 2, (B,) + size).astype(np.uint8)
    gt[gt >= K] = 255
    ld, gd = torch.from_numpy(logits).cuda(), torch.from_numpy(gt).cuda()
    poisoned = torch.full((B,) + size, 0xEE, dtype=torch.uint8, device="cuda")
    mask, C = ops.lift_argmax_hist(ld, size, gd, mask_out=poisoned)
    want = ohead.lift_argmax_spec(logits, size)
    assert np.array_equal(mask.cpu().numpy(), want)
    assert np.array_equal(C.cpu().numpy(), omet.confusion_matrix(want, gt, K))
    ops.check_status("cuda")


def test_lift_argmax_hist_odd_width_and_large_k_fallback():
    ops = _ops()
    rng = np.random.default_rng(11)
    for K, h, w, size in ((5, 4, 3, (64, 45)), (20, 4, 4, (64, 64)), (4, 16, 16, (32, 32))):
        logits = rng.standard_normal((2, K, h, w)).astype(np.float32)
        gt = rng.integers(0, K, (2,) + size).astype(np.uint8)
        mask, C = ops.lift_argmax_hist(torch.from_numpy(logits).cuda(), size, torch.from_numpy(gt).cuda())
        want = ohead.lift_argmax_spec(logits, size)
        assert np.array_equal(mask.cpu().numpy(), want)
        assert np.array_equal(C.cpu().numpy(), omet.confusion_matrix(want, gt, K))


def test_fused_hist_kernels_push_into_peer_windows():
    """world = 3 same-process windows: the fused producers carry the peer push as their tail; the reduce
    returns the sum of the three ranks' matrices (both channels), several steps in a row."""
    from ldiffusion_b200.dist import ConfusionExchange
    ops = _ops()
    K, world = 11, 3
    grp = [ConfusionExchange(K, 2, _local_group=(r, world)) for r in range(world)]
    ConfusionExchange.connect_local(grp)
    rng = np.random.default_rng(21)
    size = (256, 256)
    for step in range(6):
        want = np.zeros((2, K + 1, K), np.int64)
        for r in range(world):
            logits = rng.standard_normal((1, K, 8, 8)).astype(np.float32)
            gt = rng.integers(0, K + 1, (1,) + size).astype(np.uint8)
            inst = rng.integers(0, 50, (1,) + size).astype(np.int32)
            lut = rng.integers(0, K, (1, 50)).astype(np.uint8)
            gd = torch.from_numpy(gt).cuda()
            m0, C0 = ops.lift_argmax_hist(torch.from_numpy(logits).cuda(), size, gd, exchange=grp[r], channel=0)
            m1, C1 = ops.lut_paint_hist(torch.from_numpy(inst).cuda(), torch.from_numpy(lut).cuda(), gd, K,
                                        exchange=grp[r], channel=1)
            want[0] += omet.confusion_matrix(ohead.lift_argmax_spec(logits, size), gt, K)
            want[1] += omet.confusion_matrix(lut[0][inst], gt, K)
            assert np.array_equal(C0.cpu().numpy(), omet.confusion_matrix(m0.cpu().numpy(), gt, K))
        for r in range(world):
            assert np.array_equal(grp[r].reduce().cpu().numpy(), want), (step, r)
    ops.check_status("cuda")
    for x in grp:
        x.close()


# ---------------------------------------------------------------- whole pass: fused == separate launches

@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_fused_pass_equals_unfused_pass(dtype):
    from ldiffusion_b200 import _cabi
    from ldiffusion_b200.pipeline import HotPath, HotPathInputs, synth_inputs
    cfg = dict(batch=2, height=256, width=256, num_classes=11, num_steps=5, n_instances=40)
    host = synth_inputs(2, 256, 256, 11, 5, dtype=dtype, device="cpu", head_hw=(8, 8), n_instances=40, seed=5)
    dev = HotPathInputs(*[([t.cuda() for t in f] if isinstance(f, list) else f.cuda())
                          for f in (host.latents, host.eps, host.decoded, host.head_feat, host.inst_map,
                                    host.inst_feats, host.gt)])
    outs = []
    for fused, tissue_fused in ((True, False), (True, True), (False, False)):
        hp = HotPath(dtype=dtype, device="cuda", head_hw=(8, 8), feat_size=(16, 16), seed=5, **cfg)
        assert hp.fused and not hp.tissue_hist_fused
        hp.fused, hp.tissue_hist_fused = fused, tissue_fused
        c0 = _cabi.launch_count()
        hp.run(dev)
        assert _cabi.launch_count() - c0 == hp.launches_per_pass()
        torch.cuda.synchronize()
        outs.append({k: ([t.clone() for t in v] if isinstance(v, list) else v.clone()) for k, v in hp.results().items()})
    assert HotPath(dtype=dtype, device="cuda", head_hw=(8, 8), feat_size=(16, 16), **cfg).launches_per_pass() == 17
    for other in outs[1:]:
        for k in outs[0]:
            a, b = outs[0][k], other[k]
            for x, y in zip(a if isinstance(a, list) else [a], b if isinstance(b, list) else [b]):
                assert torch.equal(x, y), k


@pytest.mark.parametrize("nfly", [2, 3])
def test_ring_of_passes_in_flight_equals_single_passes(nfly):
    """HotPathRing: four batches through two / three slots (graph per slot, replayed on the slots' streams) give the
    same results as one HotPath run batch after batch.  Three slots run their decode tails in the throughput shape
    (one CTA per SM), selected per pass and restored: the library's setting is untouched afterwards."""
    from ldiffusion_b200.pipeline import HotPath, HotPathInputs, HotPathRing, synth_inputs
    cfg = dict(batch=2, height=256, width=256, num_classes=11, num_steps=5, n_instances=40)
    kw = dict(dtype=torch.bfloat16, device="cuda", head_hw=(8, 8), feat_size=(16, 16), seed=5, **cfg)
    devs = []
    for seed in (31, 32, 33, 34):
        host = synth_inputs(2, 256, 256, 11, 5, dtype=torch.bfloat16, device="cpu", head_hw=(8, 8), n_instances=40, seed=seed)
        devs.append(HotPathInputs(*[([t.cuda() for t in f] if isinstance(f, list) else f.cuda())
                                    for f in (host.latents, host.eps, host.decoded, host.head_feat, host.inst_map,
                                              host.inst_feats, host.gt)]))
    single = HotPath(**kw)
    want = []
    for d in devs:
        single.run(d)
        torch.cuda.synchronize()
        want.append({k: ([t.clone() for t in v] if isinstance(v, list) else v.clone()) for k, v in single.results().items()})
    from ldiffusion_b200 import _cabi
    lib = _cabi.lib()
    shape_before = lib.ldiff_tune_get(_cabi.TUNE_DECODE_TAIL_TMA)
    ring = HotPathRing(nfly, **kw)
    assert all(sl.decode_tail_shape == (HotPathRing.THROUGHPUT_DECODE_TAIL_SHAPE if nfly >= 3 else None) for sl in ring.slots)
    ring.fork()
    for i in range(nfly):                                   # warm-up outside capture (side streams, first launches)
        ring.run(i, devs[i])
    torch.cuda.synchronize()
    graphs = []
    for i in range(nfly):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=ring.stream(i)):
            ring.slot(i).run(devs[i])
        graphs.append(g)
    assert lib.ldiff_tune_get(_cabi.TUNE_DECODE_TAIL_TMA) == shape_before
    for v in ring.slot(0).results().values():               # scribble, then replay all slots concurrently
        for t_ in (v if isinstance(v, list) else [v]):
            t_.zero_()
    for i in range(nfly):
        with torch.cuda.stream(ring.stream(i)):
            graphs[i].replay()
    ring.join()
    torch.cuda.synchronize()
    for i in range(nfly):
        got = ring.slot(i).results()
        for k in want[i]:
            for x, y in zip(got[k] if isinstance(got[k], list) else [got[k]],
                            want[i][k] if isinstance(want[i][k], list) else [want[i][k]]):
                assert torch.equal(x, y), (i, k)
    for i in range(nfly, 4):                                # eager passes through the ring: slot reuse
        ring.run(i, devs[i])
    ring.join()
    torch.cuda.synchronize()
    for i in range(nfly, 4):
        got = ring.slot(i).results()
        for k in ("mask_tissue", "mask_cell", "confusion", "pixel_planes", "featcat", "latents"):
            assert torch.equal(got[k], want[i][k]), (i, k)
