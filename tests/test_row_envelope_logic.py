"""CPU restatement of the run decisions of lift_argmax_row_kernel (csrc/lift_argmax_row.cu): chunk geometry,
the lines of a row, the run-length bound and the byte-mask word building — checked against the pinned spec
(oracle/head.py::lift_argmax_spec) on small maps.  Every pixel a run claims must carry the spec's class; every
other pixel is one the kernel hands to the pinned softmax.  (Test infrastructure: numpy, exact fp32 emulation.)"""
import numpy as np
import pytest

from oracle import head as ohead
from oracle._fp import fma32, mul32, sub32
from oracle.bilinear import source_index

F = 32


def _row_form(logits, H):
    """Returns (mask with 255 where the kernel would queue the pixel for the pinned softmax, segments per chunk)."""
    B, K, h, w = logits.shape
    W = F * w
    y0, y1, hy0, hy1 = source_index(H, h)
    out = np.full((B, H, W), 254, np.uint8)
    nseg = []
    for b in range(B):
        for y in range(H):
            # vertical lerp of every class at every source column: fma(hy0, p0, hy1 * p1)
            A = fma32(hy0[y], logits[b, :, y0[y], :], mul32(hy1[y], logits[b, :, y1[y], :]))     # [K, w]
            for c in range(w + 1):
                cl, cr = max(c - 1, 0), min(c, w - 1)
                T, U = A[:, cl], A[:, cr]
                D = sub32(U, T)
                M = np.float32(max(np.abs(T).max(), np.abs(U).max()))
                gap = fma32(M, np.float32(2.0 ** -16), np.float32(1e-5))
                j0, ncol, xc = (F // 2 if c == 0 else 0), (F // 2 if c == w else F), F * c - F // 2
                words = np.zeros(F // 4, np.uint32)
                prev, r, last_unc, segs = 0, j0, False, 0
                while r < ncol:
                    l = np.float32((r + 0.5) / F)
                    v = fma32(l, D, T)
                    thr = sub32(v.max(), gap)
                    n = fma32(v, np.float32(-1), thr)
                    cand = np.signbit(n)
                    unc = (not M < 1e29) or cand.sum() != 1
                    cls, rend = 0, r + 1
                    if not unc:
                        a = int(np.argmax(cand))
                        with np.errstate(divide="ignore", invalid="ignore"):
                            q = mul32(sub32(D, D[a]), (np.float32(1) / n).astype(np.float32))
                        rmax = np.float32(max(0.0, float(np.nanmax(q))))
                        cls, rend = a, ncol
                        if rmax > 0:
                            step = mul32((np.float32(1) / rmax).astype(np.float32), np.float32(0.999755859375))
                            hi = (l + step).astype(np.float32)
                            j = int(min(np.ceil(fma32(hi, np.float32(F), np.float32(-0.5))), 1e6))
                            rend = max(r + 1, min(j, ncol))
                    if not (unc and last_unc):
                        delta = np.uint32(((cls if not unc else 0) ^ prev) * 0x01010101)
                        prev = cls if not unc else 0
                        for wd in range(F // 4):
                            sh = max(8 * r - 32 * wd, 0)
                            words[wd] ^= (np.uint32(0xFFFFFFFF << sh & 0xFFFFFFFF) if sh < 32 else np.uint32(0)) & delta
                        segs += 1
                    if unc:
                        out[b, y, xc + r] = 255
                    last_unc, r = unc, rend
                nseg.append(segs)
                row = words.view(np.uint8)                     # little-endian: byte j of the chunk
                for j in range(j0, ncol):
                    if out[b, y, xc + j] != 255:
                        out[b, y, xc + j] = row[j]
    return out, nseg


def _smooth(B, K, h, w, seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.meshgrid(np.linspace(0, 1, h), np.linspace(0, 1, w), indexing="ij")
    out = np.zeros((B, K, h, w), np.float32)
    for b in range(B):
        for k in range(K):
            fy, fx, ph = rng.uniform(0.5, 2.5, 2).tolist() + [rng.uniform(0, 6.28)]
            out[b, k] = 3 * np.sin(2 * np.pi * (fy * yy + fx * xx) + ph) + rng.normal(0, 0.05, (h, w))
    return out


CASES = {
    "iid": (lambda: np.random.default_rng(1).standard_normal((1, 11, 3, 4)).astype(np.float32), 40),
    "smooth": (lambda: _smooth(1, 7, 4, 3, 2), 48),
    "odd_k_one_column": (lambda: np.random.default_rng(2).standard_normal((1, 5, 2, 1)).astype(np.float32), 24),
    "big_magnitude": (lambda: (np.random.default_rng(5).standard_normal((1, 6, 3, 3)) * 3e4).astype(np.float32), 33),
    "within_1e-5": (lambda: (1.5 + np.random.default_rng(3).uniform(-5e-6, 5e-6, (1, 4, 2, 2))).astype(np.float32), 16),
    "ulp_pairs": (lambda: (lambda a: np.concatenate([a, np.nextafter(a, np.float32(9)), a - 1], 1))(
        np.random.default_rng(4).standard_normal((1, 1, 3, 3)).astype(np.float32) * 3), 20),
    "constant": (lambda: np.zeros((1, 3, 2, 2), np.float32), 12),
}


@pytest.mark.parametrize("case", sorted(CASES))
def test_row_form_runs_agree_with_the_spec(case):
    make, H = CASES[case]
    logits = make()
    got, nseg = _row_form(logits, H)
    want = ohead.lift_argmax_spec(logits, (H, F * logits.shape[-1]))
    assert not (got == 254).any()                              # every pixel is covered by a run or queued
    fast = got != 255
    assert np.array_equal(got[fast], want[fast])               # a run never claims a pixel the spec gives another class
    if case in ("iid", "smooth", "big_magnitude", "odd_k_one_column"):
        assert fast.mean() > 0.99                              # ... and the fast path carries the image
    if case in ("within_1e-5", "constant"):
        assert not fast.any()                                  # all near-ties: everything goes to the pinned softmax


@pytest.mark.parametrize("seed", range(8))
def test_row_form_random_maps(seed):
    """Random class counts, map sizes, magnitudes and vertical factors, with a few hand-made hazards mixed in:
    duplicated classes (exact ties along whole rows), classes an ulp apart, one class far above the rest."""
    rng = np.random.default_rng(100 + seed)
    K, h, w = int(rng.integers(2, 16)), int(rng.integers(1, 4)), int(rng.integers(1, 4))
    H = int(h * rng.integers(29, 41)) + int(rng.integers(0, 3))
    logits = (rng.standard_normal((1, K, h, w)) * float(10.0 ** rng.integers(-3, 4))).astype(np.float32)
    hazard = seed % 4
    if hazard == 1 and K >= 3:
        logits[:, 1] = logits[:, 0]                                    # an exact tie everywhere
    elif hazard == 2 and K >= 3:
        logits[:, 2] = np.nextafter(logits[:, 0], np.float32(np.inf))  # one ulp apart
    elif hazard == 3:
        logits[:, K - 1] += np.float32(50.0) * np.abs(logits).max()    # a single run per chunk
    got, nseg = _row_form(logits, H)
    want = ohead.lift_argmax_spec(logits, (H, F * w))
    assert not (got == 254).any()
    fast = got != 255
    assert np.array_equal(got[fast], want[fast]), (K, h, w, H, hazard)
    if hazard == 3:
        assert fast.all() and max(nseg) == 1
