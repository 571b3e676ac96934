"""Exactly-rounded fp32 helpers for the spec tier (test infrastructure only).

numpy has no fused multiply-add.  ``fma32`` computes ``fl32(a*b + c)`` with a
single rounding, as ``__fmaf_rn`` / x86 ``vfmadd`` do: the product of two
fp32 values is exact in fp64; the fp64 sum is turned into a round-to-odd value
with the TwoSum error term, after which the final fp64 -> fp32 rounding cannot
suffer from double rounding (Boldo & Melquiond, "Emulation of FMA and correctly
rounded sums", 2008).
"""
import numpy as np


def fma32(a, b, c):
    a, b, c = np.broadcast_arrays(np.asarray(a, np.float32), np.asarray(b, np.float32),
                                  np.asarray(c, np.float32))
    p = a.astype(np.float64) * b.astype(np.float64)          # exact
    c = c.astype(np.float64)
    s = p + c
    bb = s - p
    err = (p - (s - bb)) + (c - bb)                            # exact error of s
    bits = s.view(np.int64).copy()
    inexact = (err != 0) & np.isfinite(s)
    even = (bits & 1) == 0
    # the odd neighbour lies one ulp away from s in the direction of err
    away = (err > 0) == (s > 0)                                # |exact| > |s|
    step = np.where(away, 1, -1).astype(np.int64)
    fix = inexact & even
    # s == 0 with err != 0 cannot happen (then s would equal err exactly)
    bits = np.where(fix, bits + step, bits)
    return bits.view(np.float64).astype(np.float32)


def mul32(a, b):
    return (np.asarray(a, np.float32) * np.asarray(b, np.float32)).astype(np.float32)


def add32(a, b):
    return (np.asarray(a, np.float32) + np.asarray(b, np.float32)).astype(np.float32)


def sub32(a, b):
    return (np.asarray(a, np.float32) - np.asarray(b, np.float32)).astype(np.float32)


def div32(a, b):
    return (np.asarray(a, np.float32) / np.asarray(b, np.float32)).astype(np.float32)
