"""CPU: numpy model of the fit inside boost_blend_kernel (csrc/boost_kernels.cu) against np.polyfit as the reference calls it
(src/depthmap_generation.py:911: float32 arrays of 2^20 samples, so rcond = len(x) * eps(float32) = 0.125 and the SVD least squares on the
column-normalised Vandermonde matrix drops its second singular value for low-contrast patches).  The kernel evaluates the (possibly
rank-truncated) pseudo-inverse in closed form from five fp64 sums; this file holds the same formula in numpy and checks it in both
regimes and around the threshold."""
import warnings

import numpy as np
import pytest


def closed_form(x, y):
    xd, yd = x.astype(np.float64), y.astype(np.float64)
    n = float(x.size)
    sx, sy, sxx, sxy = xd.sum(), yd.sum(), (xd * xd).sum(), (xd * yd).sum()
    nx, nn = np.sqrt(sxx), np.sqrt(n)
    c = sx / (nx * nn)
    b0, b1 = sxy / nx, sy / nn
    rcond = n * 1.1920928955078125e-07
    c0 = c1 = (b0 + b1) / (2.0 * (1.0 + c))
    full = np.sqrt(1.0 - c) > rcond * np.sqrt(1.0 + c)
    if full:
        t = (b0 - b1) / (2.0 * (1.0 - c))
        c0, c1 = c0 + t, c1 - t
    return c0 / nx, c1 / nn, bool(full)


@pytest.mark.parametrize("spread", [0.5, 0.3, 0.2, 0.14, 0.1, 0.05, 0.01])
def test_closed_form_equals_polyfit(spread):
    rng = np.random.default_rng(int(spread * 1000))
    x = (0.5 + spread * rng.standard_normal(1 << 20)).astype(np.float32)
    y = (0.3 * x + 0.1 + 0.01 * rng.standard_normal(1 << 20)).astype(np.float32)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = np.polyfit(x, y, deg=1)
    slope, icpt, full = closed_form(x, y)
    # spread / mean = 0.25 is where sqrt(1 - c) crosses 0.125 sqrt(1 + c)
    assert full == (spread / 0.5 > 0.26) or abs(spread / 0.5 - 0.25) < 0.05
    got = np.polyval([slope, icpt], x.astype(np.float64))
    ref = np.polyval(want, x.astype(np.float64))
    assert np.abs(got - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max())          # polyfit itself runs in float32


def test_rank_truncated_fit_is_not_the_regression_line():
    """documents the quirk: for a low-contrast patch the reference's 'degree-1 fit' is far from ordinary least squares"""
    rng = np.random.default_rng(3)
    x = (0.5 + 0.02 * rng.standard_normal(1 << 20)).astype(np.float32)
    y = (2.0 * x - 0.7).astype(np.float32)
    slope, icpt, full = closed_form(x, y)
    assert not full and abs(slope - 2.0) > 0.5
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = np.polyfit(x, y, deg=1)
    assert abs(want[0] - slope) < 1e-3 and abs(want[1] - icpt) < 1e-3
