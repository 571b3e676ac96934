"""GPU parity: a-5 head + argmax and a-6 confusion matrix / metrics vs the oracle."""
import numpy as np
import pytest
import torch

from oracle import head as ohead
from oracle import metrics as omet

pytestmark = pytest.mark.gpu


def _ops():
    from ldiffusion_b200 import ops
    return ops


def _blob_labels(shape, K, seed, other_frac=0.001):
    """70% background, 16-64 px blobs of classes 1..K-1, a few 255 'other' pixels."""
    rng = np.random.default_rng(seed)
    B, H, W = shape
    out = np.zeros(shape, np.uint8)
    for b in range(B):
        for _ in range(max(1, H * W // 4000)):
            s = int(rng.integers(16, 65)); y = int(rng.integers(0, max(1, H - s))); x = int(rng.integers(0, max(1, W - s)))
            out[b, y:y + s, x:x + s] = rng.integers(1, K)
    m = rng.random(shape) < other_frac
    out[m] = 255
    return out


# ------------------------------ head ---------------------------------------

@pytest.mark.parametrize("K", [11, 7, 6, 2])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_head_logits_close(K, dtype):
    g = torch.Generator().manual_seed(K)
    feat = torch.randn(2, 256, 32, 32, generator=g).to(dtype)
    w = (torch.randn(K, 256, generator=g) / 16).to(dtype)
    bias = torch.randn(K, generator=g)
    want = ohead.conv1x1_chain(feat.float(), w.float(), bias)
    got = _ops().head_logits(feat.cuda(), w.cuda(), bias.cuda()).cpu()
    # contraction order is implementation-defined: 1e-3 relative contract (checked tighter)
    torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("K,shape,size", [(11, (2, 32, 32), (1024, 1024)), (7, (1, 32, 32), (1024, 1024)),
                                          (6, (1, 16, 16), (512, 512)), (3, (1, 5, 7), (33, 29)),
                                          (20, (1, 8, 8), (64, 64))])
def test_lift_argmax_bit_exact(K, shape, size):
    g = torch.Generator().manual_seed(100 + K)
    logits = torch.randn(shape[0], K, shape[1], shape[2], generator=g)
    want = ohead.lift_argmax_spec(logits.numpy(), size)
    got = _ops().lift_argmax(logits.cuda(), size).cpu().numpy()
    assert np.array_equal(got, want)
    chain = ohead.lift_argmax_chain(logits, size).numpy().astype(np.uint8)
    assert (chain != got).sum() == 0


def test_lift_argmax_ties_and_near_ties():
    """All-equal logits -> class 0 (first index); logits a few ulp apart go through the
    pinned softmax; small-integer logits (cell path) are exact unconditionally."""
    K = 11
    z = torch.zeros(1, K, 4, 4)
    assert int(_ops().lift_argmax(z.cuda(), (64, 64)).max()) == 0
    g = torch.Generator().manual_seed(5)
    base = torch.randn(1, 1, 8, 8, generator=g).repeat(1, K, 1, 1)
    jitter = torch.randint(-3, 4, (1, K, 8, 8), generator=g).float()
    near = base + jitter * torch.finfo(torch.float32).eps * base.abs()
    want = ohead.lift_argmax_spec(near.numpy(), (128, 128))
    got = _ops().lift_argmax(near.cuda(), (128, 128)).cpu().numpy()
    assert np.array_equal(got, want)
    ints = torch.randint(0, 3, (2, K, 16, 16), generator=g).float()
    want = ohead.lift_argmax_spec(ints.numpy(), (256, 256))
    got = _ops().lift_argmax(ints.cuda(), (256, 256)).cpu().numpy()
    assert np.array_equal(got, want)
    assert np.array_equal(got, ohead.lift_argmax_chain(ints, (256, 256)).numpy().astype(np.uint8))


def test_head_argmax_fused_path():
    """Mask is bit-exact given the kernel's own low-res logits; logits within tolerance."""
    g = torch.Generator().manual_seed(42)
    feat = torch.randn(2, 256, 32, 32, generator=g).bfloat16()
    w = (torch.randn(11, 256, generator=g) / 16).bfloat16()
    mask, logits = _ops().head_argmax(feat.cuda(), w.cuda(), None, (1024, 1024), return_logits=True)
    assert np.array_equal(mask.cpu().numpy(), ohead.lift_argmax_spec(logits.cpu().numpy(), (1024, 1024)))
    m_ref, l_ref = ohead.head_argmax_chain(feat.float(), w.float(), None, (1024, 1024))
    torch.testing.assert_close(logits.cpu(), l_ref, rtol=1e-4, atol=1e-4)
    assert (m_ref.numpy() != mask.cpu().numpy()).mean() < 1e-3     # only where logits differ by rounding


def test_cell_head_and_paint():
    K, N, H, W = 11, 300, 256, 256
    rng = np.random.default_rng(3)
    inst = np.zeros((H, W), np.int32)
    for i in range(1, N + 40):                          # ids N+1.. have no classifier row (skipped instances)
        y, x = rng.integers(0, H - 12), rng.integers(0, W - 12)
        inst[y:y + rng.integers(3, 12), x:x + rng.integers(3, 12)] = i
    ids = np.arange(1, N + 1)
    g = torch.Generator().manual_seed(8)
    feats = torch.randn(N, 256, generator=g)
    w = torch.randn(K, 256, generator=g) / 16
    b = torch.randn(K, generator=g) * 0.1
    cls_ref, logits_ref = ohead.cell_classify_chain(feats, w, b)
    lut, lo = _ops().cell_classify(feats.cuda(), w.cuda(), b.cuda(), torch.from_numpy(ids).int().cuda(),
                                   N + 40, return_logits=True)
    torch.testing.assert_close(lo.cpu(), logits_ref, rtol=1e-4, atol=1e-4)
    # decision is exact given the kernel's own logits
    p = torch.softmax(lo.cpu(), 1)[:, 1:]
    assert np.array_equal(lut.cpu().numpy()[1:N + 1], (p.argmax(1) + 1).numpy().astype(np.uint8))
    assert np.array_equal(lut.cpu().numpy()[1:N + 1], cls_ref.numpy().astype(np.uint8))
    mask = _ops().lut_paint(torch.from_numpy(inst).cuda(), lut).cpu().numpy()[0]
    want = ohead.cell_paint_spec(inst, ohead.cell_lut_spec(ids, cls_ref.numpy(), N + 40))
    assert np.array_equal(mask, want)
    # the literal painting loop of the reference on a crop (it allocates [1,K,H,W] per instance)
    crop = inst[:64, :64]
    present = [i for i in np.unique(crop) if 0 < i <= N]
    ref = ohead.cell_paint_chain(crop, present, [int(cls_ref[i - 1]) for i in present], K).numpy()
    assert np.array_equal(mask[:64, :64], ref.astype(np.uint8))


@pytest.mark.parametrize("B,N,K,Cin", [(1, 300, 11, 256), (3, 200, 11, 256), (8, 800, 11, 256), (2, 129, 6, 128),
                                        (1, 1, 2, 64), (2, 50, 16, 512)])
def test_cell_classify_tensor_core_form(B, N, K, Cin):
    """bf16 instance features take the tcgen05 form (head_tc.cu): logits within the tensor-core summation order of the
    chain on the same bf16 values, the class of every instance is the pinned rule on the kernel's OWN logits, and
    it agrees with the CUDA-core kernel (fp32 storage of the same values) wherever that kernel's top-2 gap is not a
    near-tie.  Instance counts that are not multiples of the 128-row tile exercise the zero-filled rows."""
    g = torch.Generator().manual_seed(100 + N)
    feats = torch.randn(B, N, Cin, generator=g).bfloat16()
    w = (torch.randn(K, Cin, generator=g) / 16).bfloat16()
    b = torch.randn(K, generator=g) * 0.1
    ids = torch.arange(1, N + 1, dtype=torch.int32)
    lut, lo = _ops().cell_classify(feats.cuda(), w.cuda(), b.cuda(), ids.cuda(), N + 1, return_logits=True)
    lut_s, lo_s = _ops().cell_classify(feats.float().cuda(), w.float().cuda(), b.cuda(), ids.cuda(), N + 1,
                                       return_logits=True)
    for i in range(B):
        cls_ref, logits_ref = ohead.cell_classify_chain(feats[i].float(), w.float(), b)
        torch.testing.assert_close(lo[i].cpu(), logits_ref, rtol=1e-4, atol=1e-4)
        if K > 1:
            p = torch.softmax(lo[i].cpu(), 1)[:, 1:]
            assert np.array_equal(lut[i].cpu().numpy()[1:], (p.argmax(1) + 1).numpy().astype(np.uint8))
        else:
            assert not lut[i].any()
        diff = (lut[i] != lut_s[i]).cpu().numpy()[1:]
        if diff.any():                                   # only near-ties of the CUDA-core kernel's logits may flip
            top2 = torch.topk(lo_s[i].cpu()[:, 1:], 2, dim=1).values
            assert ((top2[:, 0] - top2[:, 1])[torch.from_numpy(diff)] < 1e-4).all()
        assert lut[i, 0] == 0
    _ops().check_status(torch.device("cuda"))


def test_cell_classify_tensor_core_id_range_error():
    feats = torch.randn(1, 130, 256).bfloat16().cuda()
    w = (torch.randn(11, 256) / 16).bfloat16().cuda()
    ids = torch.arange(1, 131, dtype=torch.int32).cuda()
    _ops().cell_classify(feats, w, None, ids, 100)       # ids 100..130 do not fit a 100-entry LUT
    with pytest.raises(RuntimeError):
        _ops().check_status(torch.device("cuda"))


def test_lut_paint_range_error():
    inst = torch.tensor([[0, 1, 5, 2]], dtype=torch.int32).repeat(4, 4).cuda()
    lut = torch.tensor([0, 3, 4], dtype=torch.uint8).cuda()
    _ops().lut_paint(inst, lut)
    with pytest.raises(RuntimeError):
        _ops().check_status(inst.device)


# ------------------------------ metrics -------------------------------------

@pytest.mark.parametrize("K,shape", [(11, (1, 512, 512)), (7, (2, 256, 256)), (6, (1, 64, 64)),
                                     (11, (1, 37, 53)), (2, (1, 16, 16)), (11, (8, 1024, 1024))])
def test_confusion_matrix_bit_exact(K, shape):
    gt = _blob_labels(shape, K, 1)
    pred = _blob_labels(shape, K, 2, other_frac=0)
    want = omet.confusion_matrix(pred, gt, K)
    got = _ops().confusion_hist(torch.from_numpy(pred).cuda().reshape(-1), torch.from_numpy(gt).cuda().reshape(-1), K)
    _ops().check_status("cuda")
    assert np.array_equal(got.cpu().numpy(), want)
    assert int(got.sum()) == pred.size


def test_confusion_random_labels_and_accumulate():
    """Worst case for run merging (i.i.d. labels) and accumulation into the same buffer."""
    K = 11
    rng = np.random.default_rng(4)
    C = None
    want = np.zeros((K + 1, K), np.int64)
    for i in range(3):
        pred = rng.integers(0, K, (1, 300, 301)).astype(np.uint8)
        gt = rng.integers(0, 14, (1, 300, 301)).astype(np.uint8)
        want += omet.confusion_matrix(pred, gt, K)
        C = _ops().confusion_hist(torch.from_numpy(pred).cuda().reshape(-1),
                                  torch.from_numpy(gt).cuda().reshape(-1), K, out=C)
    assert np.array_equal(C.cpu().numpy(), want)


def test_confusion_batched_per_image():
    """One launch, one matrix per image (evaluate.py averages per-image scores)."""
    K = 11
    gt = _blob_labels((5, 256, 320), K, 31)
    pred = _blob_labels((5, 256, 320), K, 32, other_frac=0)
    got = _ops().confusion_hist_batched(torch.from_numpy(pred).cuda(), torch.from_numpy(gt).cuda(), K).cpu().numpy()
    for i in range(5):
        assert np.array_equal(got[i], omet.confusion_matrix(pred[i], gt[i], K))
    odd = _ops().confusion_hist_batched(torch.from_numpy(pred[:, :37, :53].copy()).cuda(),
                                        torch.from_numpy(gt[:, :37, :53].copy()).cuda(), K).cpu().numpy()
    for i in range(5):
        assert np.array_equal(odd[i], omet.confusion_matrix(pred[i, :37, :53], gt[i, :37, :53], K))


@pytest.mark.parametrize("K", [16, 40, 128])
def test_confusion_large_k(K):
    rng = np.random.default_rng(K)
    pred = rng.integers(0, K, (1, 200, 200)).astype(np.uint8)
    gt = rng.integers(0, min(K + 3, 256), (1, 200, 200)).astype(np.uint8)
    got = _ops().confusion_hist(torch.from_numpy(pred).cuda().reshape(-1), torch.from_numpy(gt).cuda().reshape(-1), K)
    assert np.array_equal(got.cpu().numpy(), omet.confusion_matrix(pred, gt, K))


def test_confusion_gt_lut_fusion():
    """dataset.py:20-32 gray-level -> class LUT applied inside the histogram pass."""
    K = 11
    lut = np.zeros(256, np.uint8)
    for k, v in {0: 0, 25: 1, 50: 2, 75: 3, 100: 4, 125: 5, 150: 6, 175: 7, 200: 8, 225: 9, 250: 10}.items():
        lut[k] = v
    cls = _blob_labels((1, 256, 256), K, 6, other_frac=0)
    raw = (cls.astype(np.int32) * 25).astype(np.uint8)
    pred = _blob_labels((1, 256, 256), K, 7, other_frac=0)
    got = _ops().confusion_hist(torch.from_numpy(pred).cuda().reshape(-1), torch.from_numpy(raw).cuda().reshape(-1),
                                K, gt_lut=torch.from_numpy(lut).cuda())
    assert np.array_equal(got.cpu().numpy(), omet.confusion_matrix(pred, cls, K))


def test_pred_out_of_range_raises_like_one_hot():
    from ldiffusion_b200 import micro_dice
    pred = torch.full((1, 32, 32), 11, dtype=torch.uint8).cuda()
    gt = torch.zeros((32, 32), dtype=torch.int64).cuda()
    with pytest.raises(RuntimeError, match="Class values must be smaller"):
        micro_dice(pred, gt, 11)


@pytest.mark.parametrize("K,hw", [(11, 512), (7, 256), (6, 128)])
def test_dropin_metric_functions_match_reference_chain(K, hw):
    """Same signatures as utils.py / evaluate.py: one-hot float preds + int64 gt."""
    import ldiffusion_b200 as L
    gt = torch.from_numpy(_blob_labels((1, hw, hw), K, 21)).long()[0]
    pred = torch.from_numpy(_blob_labels((1, hw, hw), K, 22, other_frac=0)).long()
    onehot = torch.nn.functional.one_hot(pred, K).permute(0, 3, 1, 2).float()
    d_ref, a_ref = omet.micro_dice_chain(onehot, gt, K)
    d, a = L.micro_dice(onehot.cuda(), gt.cuda(), K)
    assert torch.equal(d.cpu(), d_ref) and torch.equal(a.cpu(), a_ref)
    assert L.mean_iou_and_per_class(onehot.cuda(), gt.cuda(), K) == omet.mean_iou_and_per_class_chain(onehot, gt, K)
    assert L.pixel_accuracy(onehot.cuda(), gt.cuda(), K) == omet.pixel_accuracy_chain(onehot, gt, K)
    for ib in (False, True):
        assert L.frequency_weighted_iou(onehot.cuda(), gt.cuda(), K, ib) == \
            omet.frequency_weighted_iou_chain(onehot, gt, K, ib)


def test_evaluate_folder(tmp_path):
    """evaluate.py:48-126 end to end on PNG folders."""
    from PIL import Image
    import ldiffusion_b200 as L
    K = 11
    preds = [_blob_labels((1, 128, 160), K, 60 + i, other_frac=0)[0] for i in range(3)]
    gts = [_blob_labels((1, 128, 160), K, 70 + i)[0] for i in range(3)]
    (tmp_path / "p").mkdir(); (tmp_path / "g").mkdir()
    for i, (p, g) in enumerate(zip(preds, gts)):
        Image.fromarray(p).save(tmp_path / "p" / f"{i}.png"); Image.fromarray(g).save(tmp_path / "g" / f"{i}.png")
    s = L.evaluate(str(tmp_path / "p"), str(tmp_path / "g"), K, str(tmp_path / "out"))
    want = omet.evaluate_images_chain(preds, gts, K)
    for k in ("mean_dice", "mean_iou", "mean_pa", "mean_fwiou"):
        assert s[k] == want[k], k
    for k in ("per_class_dice", "per_class_iou", "per_class_pa"):
        assert np.array_equal(np.asarray(s[k]), np.asarray(want[k])), k
    assert len(list((tmp_path / "out").glob("metrics_*.txt"))) == 1


def test_argmax_channels_first_max():
    g = torch.Generator().manual_seed(3)
    x = torch.randint(0, 3, (2, 7, 33, 31), generator=g).float()       # many ties
    assert torch.equal(_ops().argmax_channels(x.cuda()).cpu().long(), torch.argmax(x, 1))
    xb = torch.randn(1, 11, 64, 64, generator=g).bfloat16()
    assert torch.equal(_ops().argmax_channels(xb.cuda()).cpu().long(), torch.argmax(xb.float(), 1))


def test_nnunet_online_counts_golden():
    """Widening N4: nnUNetTrainer.validation_step's tp/fp/fn (vendored get_tp_fp_fn_tn, golden
    vectors) from one argmax + one histogram pass."""
    import os
    from ldiffusion_b200.metrics import online_tp_fp_fn
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "nnunet_counts.npz"))
    K = int(z["K"])
    output, target = torch.from_numpy(z["output"]).cuda(), torch.from_numpy(z["target"])
    plain = target.clone(); plain[plain == K] = 0
    got = online_tp_fp_fn(output, plain.cuda())
    assert np.array_equal(np.stack(got), z["plain"])
    ign = target.clone(); ign[ign == K] = 255            # ignore label -> outside [0,K)
    got = online_tp_fp_fn(output, ign.cuda())
    assert np.array_equal(np.stack(got), z["ignore"])
