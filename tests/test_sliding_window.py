"""Widening N2: nnU-Net sliding-window helpers (CPU, golden from the vendored functions) and the
accumulate / TTA-merge / export kernels (GPU, against the literal eager fp16 chain)."""
import ast
import os

import numpy as np
import pytest
import torch

from oracle import sliding_window as osw

Z = np.load(os.path.join(os.path.dirname(__file__), "golden", "nnunet_sliding.npz"))


def test_steps_match_vendored_nnunet():
    from ldiffusion_b200.sliding_window import compute_steps_for_sliding_window
    cases, steps = ast.literal_eval(str(Z["cases"])), ast.literal_eval(str(Z["steps"]))
    for (img, tile, frac), want in zip(cases, steps):
        assert compute_steps_for_sliding_window(img, tile, frac) == want


def test_gaussian_matches_vendored_nnunet():
    from ldiffusion_b200.sliding_window import compute_gaussian
    for ts in ((64, 64), (48, 80)):
        g = compute_gaussian(ts, sigma_scale=1. / 8, value_scaling_factor=10, device="cpu")
        assert g.dtype == torch.float16 and np.array_equal(g.numpy(), Z[f"gauss_{ts[0]}x{ts[1]}"])
    g = compute_gaussian((512, 512), sigma_scale=1. / 8, value_scaling_factor=10, device="cpu")
    assert np.array_equal(g.numpy()[::8, ::8], Z["gauss_512x512"]) and float(g.min()) > 0


def _net(K, seed=0):
    torch.manual_seed(seed)
    return torch.nn.Conv2d(3, K, 3, padding=1)


@pytest.mark.gpu
@pytest.mark.parametrize("K,H,W,tile,use_gauss", [(7, 160, 144, (64, 64), True), (5, 96, 96, (48, 80), True),
                                                  (3, 64, 64, (64, 64), False)])
def test_accumulate_and_export_match_eager_chain(K, H, W, tile, use_gauss):
    from ldiffusion_b200.sliding_window import (SlidingWindowAccumulator, compute_gaussian,
                                                compute_steps_for_sliding_window)
    g = torch.Generator().manual_seed(K)
    steps = compute_steps_for_sliding_window((H, W), tile, 0.5)
    slicers = [(slice(None), slice(sy, sy + tile[0]), slice(sx, sx + tile[1])) for sy in steps[0] for sx in steps[1]]
    preds = [(torch.randn(K, *tile, generator=g) * 3).half() for _ in slicers]
    gauss = compute_gaussian(tile, 1. / 8, 10, device="cpu") if use_gauss else None
    want_logits, _ = osw.accumulate_chain(preds, slicers, gauss, K, (H, W))
    want_seg = osw.export_chain(want_logits)
    sw = SlidingWindowAccumulator(K, (H, W), tile, use_gauss, "cuda")
    if use_gauss:
        assert torch.equal(sw.gaussian.cpu(), gauss)
    for p, sl in zip(preds, slicers):
        sw.add(p.cuda(), sl[1].start, sl[2].start)
    seg, logits = sw.finalize(return_logits=True)
    assert torch.equal(logits.cpu(), want_logits)                         # fp16, bit for bit
    assert torch.equal(seg.cpu().long(), want_seg)


@pytest.mark.gpu
def test_tta_merge_matches_eager_chain():
    from ldiffusion_b200.sliding_window import tta_merge
    K = 4
    net = _net(K)
    x = torch.randn(1, 3, 40, 56)
    want = osw.tta_merge_chain(net, x, (0, 1))[0]
    with torch.no_grad():
        preds = [net(x)[0].half(), net(torch.flip(x, (2,)))[0].half(), net(torch.flip(x, (3,)))[0].half(),
                 net(torch.flip(x, (2, 3)))[0].half()]
    got = tta_merge([p.cuda().contiguous() for p in preds], [0, 1, 2, 3])
    assert torch.equal(got.cpu(), want.detach())


@pytest.mark.gpu
def test_export_ties_in_half_precision():
    """fp16 logits tie often; equal heads -> first index, as torch.argmax(softmax) gives."""
    from ldiffusion_b200.sliding_window import SlidingWindowAccumulator
    K, H, W = 6, 32, 32
    g = torch.Generator().manual_seed(1)
    p = (torch.randint(-2, 3, (K, H, W), generator=g).float() * 0.5).half()       # many exact ties
    sw = SlidingWindowAccumulator(K, (H, W), (H, W), False, "cuda")
    sw.add(p.cuda(), 0, 0)
    seg = sw.finalize()
    want = osw.export_chain(osw.accumulate_chain([p], [(slice(None), slice(0, H), slice(0, W))], None, K, (H, W))[0])
    assert torch.equal(seg.cpu().long(), want)


@pytest.mark.gpu
def test_inf_raises_like_the_reference():
    from ldiffusion_b200 import ops
    from ldiffusion_b200.sliding_window import SlidingWindowAccumulator
    sw = SlidingWindowAccumulator(2, (16, 16), (16, 16), False, "cuda")
    p = torch.full((2, 16, 16), 60000.0).half().cuda()
    sw.add(p, 0, 0); sw.add(p, 0, 0)                                       # 120000 overflows fp16
    sw.finalize()
    with pytest.raises(RuntimeError, match="Encountered inf"):
        ops.check_status("cuda")


@pytest.mark.gpu
def test_predict_sliding_window_end_to_end():
    from ldiffusion_b200.sliding_window import compute_gaussian, compute_steps_for_sliding_window, predict_sliding_window
    K, H, W, tile = 5, 150, 130, (64, 64)
    net = _net(K, 3).cuda()
    data = torch.randn(3, H, W, device="cuda")
    seg, logits = predict_sliding_window(net, data, tile, K, return_logits=True)
    # the eager chain on the CPU with the SAME per-tile network outputs (taken from the GPU net)
    steps = compute_steps_for_sliding_window((H, W), tile, 0.5)
    gauss = compute_gaussian(tile, 1. / 8, 10, device="cpu")
    preds, slicers = [], []
    with torch.no_grad():
        for sy in steps[0]:
            for sx in steps[1]:
                x = data[None, :, sy:sy + 64, sx:sx + 64]
                pr = net(x).half()
                for axes in ((2,), (3,), (2, 3)):
                    pr += torch.flip(net(torch.flip(x, axes)).half(), axes)
                pr /= 4
                preds.append(pr[0].cpu()); slicers.append((slice(None), slice(sy, sy + 64), slice(sx, sx + 64)))
    want_logits, _ = osw.accumulate_chain(preds, slicers, gauss, K, (H, W))
    assert torch.equal(logits.cpu(), want_logits)
    assert torch.equal(seg.cpu().long(), osw.export_chain(want_logits))
