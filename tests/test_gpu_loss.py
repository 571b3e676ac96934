"""GPU parity of the widening row N3: pixel-contrastive InfoNCE forward and backward against the
reference's per-triple loop (oracle/loss.py) with injected index sets."""
import numpy as np
import pytest
import torch

from oracle.loss import contrastive_loss_chain

pytestmark = pytest.mark.gpu


def _labels(B, h, w, K, seed):
    g = torch.Generator().manual_seed(seed)
    lab = torch.zeros(B, 1, h, w, dtype=torch.uint8)
    for b in range(B):
        for _ in range(6):
            y, x = int(torch.randint(0, h - 16, (1,), generator=g)), int(torch.randint(0, w - 16, (1,), generator=g))
            lab[b, 0, y:y + 16, x:x + 16] = int(torch.randint(1, K, (1,), generator=g))
    return lab


@pytest.mark.parametrize("n,N,T", [(5, 1024, 0.5), (1, 64, 0.1), (11, 200, 1.0)])
def test_infonce_forward_backward_match_reference_loop(n, N, T):
    from ldiffusion_b200.loss import pixel_contrastive_loss, sample_contrastive_pairs
    B, h, w = 2, 64, 64
    g = torch.Generator().manual_seed(n)
    labels = _labels(B, h, w, 7, n)
    pairs = sample_contrastive_pairs(labels, num_negatives=N, generator=g)
    assert pairs is not None and pairs[3].shape[1] == N
    feats = torch.randn(B, n, h, w, generator=g)
    f_ref = feats.clone().requires_grad_(True)
    want = contrastive_loss_chain(f_ref, pairs, T)
    want.backward()
    f_dev = feats.cuda().requires_grad_(True)
    got = pixel_contrastive_loss(f_dev, labels, temperature=T, num_negatives=N, pairs=pairs)
    got.backward()
    torch.testing.assert_close(got.detach().cpu(), want.detach(), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(f_dev.grad.cpu(), f_ref.grad, rtol=1e-4, atol=1e-7)
    # upstream gradient scaling
    f2 = feats.cuda().requires_grad_(True)
    (3.0 * pixel_contrastive_loss(f2, labels, temperature=T, num_negatives=N, pairs=pairs)).backward()
    # two runs of the atomics-based backward: same sums in a different order
    torch.testing.assert_close(f2.grad, 3.0 * f_dev.grad, rtol=1e-4, atol=3e-7)


def test_infonce_sampling_follows_reference_rules():
    from ldiffusion_b200.loss import pixel_contrastive_loss, sample_contrastive_pairs
    labels = _labels(1, 64, 64, 5, 3)
    pb, pa, pq, neg = sample_contrastive_pairs(labels, 1024, torch.Generator().manual_seed(0))
    lab = labels.reshape(-1)
    assert (lab[pa.long()] == lab[pq.long()]).all() and (pa != pq).all()        # positive: same label, other pixel
    assert (lab[neg.long()] != lab[pa.long()][:, None]).all()                   # negatives: other labels
    for lbl in torch.unique(lab):                                               # 1 % of each class as anchors
        npos = int((lab == lbl).sum())
        if npos > 1 and int((lab != lbl).sum()) > 1024:
            assert int((lab[pa.long()] == lbl).sum()) == max(1, int(0.01 * npos))
    # a single-class map has no negatives: the reference returns a zero that requires grad
    flat = torch.zeros(1, 1, 64, 64, dtype=torch.uint8)
    for sampler in ("host", "device"):
        f = torch.randn(1, 5, 64, 64).cuda().requires_grad_(True)
        z = pixel_contrastive_loss(f, flat, sampler=sampler)
        assert float(z.detach()) == 0.0 and z.requires_grad
        if sampler == "device":
            z.backward()
            assert float(f.grad.abs().sum()) == 0.0


@pytest.mark.parametrize("hw,N", [((64, 64), 1024), ((32, 48), 100), ((128, 128), 1024)])
def test_device_sampler_follows_reference_rules(hw, N):
    """ldiff_infonce_sample against the rules of loss.py:64-87 (the draws themselves cannot be compared:
    the reference's are unseeded torch.randperm calls)."""
    from ldiffusion_b200 import ops
    from ldiffusion_b200.loss import sample_contrastive_pairs_device
    h, w = hw
    B = 3
    labels = torch.zeros(B, 1, h, w, dtype=torch.uint8)
    g = torch.Generator().manual_seed(h)
    for b in range(B):
        for k in range(1, 7):
            y, x = int(torch.randint(0, h - 12, (1,), generator=g)), int(torch.randint(0, w - 12, (1,), generator=g))
            labels[b, 0, y:y + 12, x:x + 12] = k
    labels[2, 0, 0, :2] = 9                                       # a 2-pixel class: max(1, 0) = 1 anchor
    labels[2, 0, 1, 0] = 10                                       # a 1-pixel class: skipped
    pb, pa, pq, neg, nv = [t.cpu() for t in sample_contrastive_pairs_device(labels.cuda(), N, seed=7, offset=3)]
    ops.check_status("cuda")
    cap = pb.numel() // B
    flat = labels.reshape(B, -1)
    for b in range(B):
        lab = flat[b]
        sl = slice(b * cap, b * cap + int(nv[b]))
        assert (pb[sl] == b).all() and (pb[b * cap + int(nv[b]):(b + 1) * cap] == -1).all()
        a, q, ng = pa[sl].long(), pq[sl].long(), neg[sl].long()
        assert (lab[a] == lab[q]).all() and (a != q).all()
        assert (lab[ng] != lab[a][:, None]).all()
        assert all(len(set(row.tolist())) == N for row in ng)                  # negatives are distinct
        want = 0
        for lbl in torch.unique(lab):                                          # ascending, as the slots are
            npos = int((lab == lbl).sum())
            if npos > 1 and int((lab != lbl).sum()) > N:
                k = max(1, int(0.01 * npos))
                cls = a[want:want + k]
                assert (lab[cls] == lbl).all() and len(set(cls.tolist())) == k  # distinct anchors of this class
                want += k
        assert want == int(nv[b])
    # deterministic in (seed, offset); a new offset gives new draws
    again = [t.cpu() for t in sample_contrastive_pairs_device(labels.cuda(), N, seed=7, offset=3)]
    assert all(torch.equal(x, y) for x, y in zip((pb, pa, pq, neg, nv), again))
    other = sample_contrastive_pairs_device(labels.cuda(), N, seed=7, offset=4)[3].cpu()
    assert not torch.equal(other, neg)


def test_device_sampler_is_uniform_enough():
    """Every pixel outside the class is drawn as a negative with probability N/M: pooled over many
    (seed, anchor) draws the per-pixel counts stay within 6 sigma of the binomial expectation; the
    anchors and positives cover their class."""
    from ldiffusion_b200.loss import sample_contrastive_pairs_device
    labels = torch.zeros(1, 1, 64, 64, dtype=torch.uint8)
    labels[0, 0, :32, :32] = 1                                    # class 1: 1024 px -> 10 anchors; class 0: 3072 -> 30
    lab = labels.cuda()
    counts = torch.zeros(4096, dtype=torch.int64)
    anchors = torch.zeros(4096, dtype=torch.int64)
    pos = torch.zeros(4096, dtype=torch.int64)
    draws = 0
    for off in range(60):
        pb, pa, pq, neg, nv = sample_contrastive_pairs_device(lab, 512, seed=1, offset=off)
        n1 = 10                                                   # class 0 slots come first (30), then class 1
        sel = neg[30:30 + n1].cpu().long().reshape(-1)            # negatives of class-1 anchors: the 3072 class-0 px
        counts += torch.bincount(sel, minlength=4096)
        anchors += torch.bincount(pa[:40].cpu().long(), minlength=4096)
        pos += torch.bincount(pq[:40].cpu().long(), minlength=4096)
        draws += n1
    flat = labels.reshape(-1)
    c0 = counts[flat == 0].double()
    p = 512 / 3072
    mu, sd = draws * p, (draws * p * (1 - p)) ** 0.5
    assert counts[flat == 1].sum() == 0
    assert float((c0 - mu).abs().max()) < 6 * sd, (float(c0.min()), float(c0.max()), mu, sd)
    assert abs(float(c0.mean()) - mu) < 1e-9                      # every draw lands somewhere in the pool
    assert int((anchors > 0).sum()) > 1500 and int((pos > 0).sum()) > 1500   # 2400 draws spread over 4096 px


def test_infonce_device_sampled_loss_matches_reference_loop():
    """End to end with the device sampler: the loss and gradient over the sampled (padded) triple list
    equal the reference loop run on the same triples."""
    from ldiffusion_b200.loss import pixel_contrastive_loss, sample_contrastive_pairs_device
    labels = _labels(2, 64, 64, 6, 11)
    feats = torch.randn(2, 5, 64, 64, generator=torch.Generator().manual_seed(2))
    pb, pa, pq, neg, nv = [t.cpu() for t in sample_contrastive_pairs_device(labels.cuda(), 1024, seed=5, offset=9)]
    keep = pb >= 0
    pairs = (pb[keep], pa[keep], pq[keep], neg[keep])
    f_ref = feats.clone().requires_grad_(True)
    want = contrastive_loss_chain(f_ref, pairs, 0.5)
    want.backward()
    f_dev = feats.cuda().requires_grad_(True)
    got = pixel_contrastive_loss(f_dev, labels.cuda(), temperature=0.5, num_negatives=1024, seed=5, offset=9)
    got.backward()
    torch.testing.assert_close(got.detach().cpu(), want.detach(), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(f_dev.grad.cpu(), f_ref.grad, rtol=1e-4, atol=1e-7)


def test_device_sampler_label_range_raises():
    from ldiffusion_b200 import ops
    from ldiffusion_b200.loss import sample_contrastive_pairs_device
    labels = torch.zeros(1, 1, 16, 16, dtype=torch.uint8)
    labels[0, 0, 0, 0] = 40
    sample_contrastive_pairs_device(labels.cuda(), 8)
    with pytest.raises(RuntimeError, match="label value"):
        ops.check_status("cuda")
