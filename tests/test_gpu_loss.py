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
    torch.testing.assert_close(f2.grad, 3.0 * f_dev.grad, rtol=1e-5, atol=1e-8)


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
    z = pixel_contrastive_loss(torch.randn(1, 5, 64, 64).cuda(), flat)
    assert float(z) == 0.0 and z.requires_grad
