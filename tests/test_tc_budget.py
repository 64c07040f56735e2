"""CPU: the error budget of the tensor sweeps' integer items, checked numerically.

The tensor sweeps evaluate 2<X, Y> of two stacked windows (9 frames x 12 bins) from 24-bit quantised features split into
byte limbs h, l1, l2:  T = acc0 * 2^e0 + ((acc1 * 256 + acc2) >> (16 - e0))  with  acc0 = h.h, acc1 = h.l1 + l1.h,
acc2 = l1.l1 + h.l2 + l2.h  (the products 2^8 (l1.l2 + l2.l1) + l2.l2 are dropped).  The design claims, in fixed-point
units 2^(fx_exp - 23):  |T - 2<X, Y>| <= 1.4 * 2^(e0/2) (quantisation) + 0.84 * 2^e0 (dropped products) + 1 (floor), and
enables the sweeps only for e0 <= 6.  This test restates the arithmetic in numpy (the same integer formula the kernels
use) and checks the claim on HPCP-like windows, on adversarial limb patterns and across feature scales, together with the
exponent selection of acoss_set_tracks (api.cu: finish_tracks)."""
import math

import numpy as np


def exponents(frames):
    """(fx_exp, q_exp, e0) as acoss_set_tracks derives them from the track set (api.cu: finish_tracks)."""
    gmax2 = float((frames.astype(np.float32) ** 2).sum(axis=1).max())
    _, fx_exp = math.frexp(2.0 * gmax2 * 1.001)
    _, e = math.frexp(float(frames.max()) * 1.000001)
    q = 24 - e
    return fx_exp, q, 56 - 2 * q - fx_exp


def tc_T(X, Y, q, e0):
    """The kernels' integer item part for stacked windows X, Y (108 floats each): three limb-product accumulators."""
    xq = np.rint(X.astype(np.float64) * 2.0 ** q).astype(np.int64)
    yq = np.rint(Y.astype(np.float64) * 2.0 ** q).astype(np.int64)
    assert xq.max() < 2 ** 24 and yq.max() < 2 ** 24 and xq.min() >= 0 and yq.min() >= 0
    xh, xl1, xl2 = xq >> 16, (xq >> 8) & 255, xq & 255
    yh, yl1, yl2 = yq >> 16, (yq >> 8) & 255, yq & 255
    a0 = int((xh * yh).sum())
    a1 = int((xh * yl1 + xl1 * yh).sum())
    a2 = int((xl1 * yl1 + xh * yl2 + xl2 * yh).sum())
    assert a1 * 256 + a2 < 2 ** 32 and a0 < 2 ** 31                        # what tc_item relies on
    return a0 * (1 << e0) + ((a1 * 256 + a2) >> (16 - e0))


def bound(e0):
    return 1.4 * 2 ** (e0 / 2) + 0.84 * 2 ** e0 + 1


def hpcp_like(rng, n):
    f = rng.random((n, 12)).astype(np.float32) ** 3                      # a few strong bins, many small ones
    f /= f.max(axis=1, keepdims=True)                                     # essentia HPCP: every frame normalised to max 1
    return f.astype(np.float32)


def check_set(frames, rng, npairs=400):
    fx_exp, q, e0 = exponents(frames)
    assert 0 <= e0 <= 8
    unit = 2.0 ** (fx_exp - 23)
    n = len(frames) - 9
    worst = 0.0
    for _ in range(npairs):
        i, j = rng.integers(0, n, size=2)
        X, Y = frames[i:i + 9].ravel(), frames[j:j + 9].ravel()
        exact = 2.0 * float((X.astype(np.float64) * Y.astype(np.float64)).sum()) / unit
        worst = max(worst, abs(tc_T(X, Y, q, e0) - exact))
    return e0, worst


def test_budget_on_hpcp_like_windows():
    rng = np.random.default_rng(1)
    frames = hpcp_like(rng, 600)
    e0, worst = check_set(frames, rng)
    assert e0 in (5, 6)                                                   # max feature 1, largest squared norm in [4, 16)
    assert worst <= bound(e0)


def test_budget_adversarial_limbs():
    """Every low limb at its maximum (the dropped products at their bound) and at zero, mixed signs of the rounding."""
    rng = np.random.default_rng(2)
    q = 23
    base = (np.full((40, 12), 0x7fffff, dtype=np.int64) - rng.integers(0, 2, size=(40, 12)) * 0x010000)   # h large, l1 = l2 = 255
    frames = (base / 2.0 ** q).astype(np.float64)
    frames[0, 0] = 1.0                                                    # pins the exponent of the largest feature
    fx_exp, q2, e0 = exponents(frames.astype(np.float32))
    # float32 cannot hold these 23-bit patterns exactly: evaluate the budget with float64 "features" (the claim is about the
    # integer arithmetic; float32 inputs are a subset)
    unit = 2.0 ** (fx_exp - 23)
    worst = 0.0
    for i in range(30):
        X, Y = frames[i:i + 9].ravel(), frames[i + 1:i + 10].ravel()
        exact = 2.0 * float((X * Y).sum()) / unit
        worst = max(worst, abs(tc_T(X, Y, q2, e0) - exact))
    assert worst <= bound(e0)
    assert worst > 0.25 * 0.84 * 2 ** e0                                   # ... and the dropped products really are that large here


def test_budget_across_scales_and_exponent_gate():
    rng = np.random.default_rng(3)
    base = hpcp_like(rng, 300)
    for scale in (1.0 / 4096, 0.37, 1.0, 3.0, 1000.0):
        frames = (base * np.float32(scale)).astype(np.float32)
        e0, worst = check_set(frames, rng, 150)
        assert worst <= bound(e0)
    # single dominant bin per frame: largest squared norm ~ (largest feature)^2  =>  e0 = 7 or 8: the library keeps the FFMA2
    # sweeps there (api.cu enables the tensor sweeps for e0 <= 6 only), because 0.84 * 2^e0 alone would pass EPS = 128
    spiky = np.full((100, 12), 1e-3, dtype=np.float32)
    spiky[np.arange(100), rng.integers(0, 12, size=100)] = 1.0
    _, _, e0 = exponents(spiky)
    assert e0 >= 7
    assert bound(7) + 17 + 1 > 128                                        # (17: the oracle's float32 roundings, 1: the two norm roundings)
    assert bound(6) + 17 + 1 <= 84
