"""Accuracy model of the device FM discriminator (sdr_b200/csrc/demod.cuh: hs_atan2f_dev, fm_phase) checked on the CPU.
The device code replaces GHC's class-default atan2 (IEEE division + libm atanf, restated in oracle/sdr_oracle.c
hs_atan2f) by reciprocal + degree-7 minimax polynomial + octant unfolding.  This test re-evaluates exactly that operation
chain in float32 numpy -- the polynomial coefficients are parsed out of demod.cuh so the two cannot drift apart -- over
noise at several scales, the u8 sample grid and a soup of signed zeros / infinities / denormals / huge values, against the
oracle's fmDemod: same value table, and never more than 1e-6 apart (the bar is 1e-5; the GPU parity tests hold the real
kernel to it)."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
f32 = np.float32


def coefficients():
    src = open(os.path.join(ROOT, "sdr_b200", "csrc", "demod.cuh")).read()
    body = src[src.index("float hs_atan2f_dev"):src.index("return copysignf")]
    first = re.search(r"float p = (-?[0-9.e-]+)f;", body).group(1)
    rest = re.findall(r"__fmaf_rn\(p, u, (-?[0-9.e-]+)f\)", body)
    assert len(rest) == 7
    return [f32(first)] + [f32(c) for c in rest]   # highest degree first


def fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + np.float64(c)).astype(f32)


def device_atan2(y, x, coef):
    ax, ay = np.abs(x), np.abs(y)
    mx, mn = np.maximum(ax, ay), np.minimum(ax, ay)          # NaN-propagating, like max.NaN / min.NaN
    sc = np.where(mx > f32(2.0 ** 100), f32(2.0 ** -64), f32(1)).astype(f32)
    with np.errstate(all="ignore"):
        t = ((mn * sc).astype(f32) / (mx * sc).astype(f32)).astype(f32)
        u = (t * t).astype(f32)
        p = np.full_like(t, coef[0])
        for c in coef[1:]:
            p = fma(p, u, c)
        p = (p * t).astype(f32)
        pi = f32(3.14159265358979323846)
        half_pi = f32(1.57079632679489661923)
        p = np.where(ay > ax, (half_pi - p).astype(f32), p)
        p = np.where(x < 0, (pi - p).astype(f32), p)
    return np.copysign(p, y).astype(f32)


def device_fm_demod(x, coef):
    s, l = x[1:], x[:-1]
    sx, sy, lx, ly = (v.astype(f32) for v in (s.real, s.imag, l.real, l.imag))
    nli = -ly
    with np.errstate(all="ignore"):
        re_ = ((sx * lx).astype(f32) - (sy * nli).astype(f32)).astype(f32)
        im_ = ((sx * nli).astype(f32) + (sy * lx).astype(f32)).astype(f32)
    return np.where((re_ == 0) & (im_ == 0), f32(0), device_atan2(im_, re_, coef))


def compare(port, x, coef):
    want = port.fm_demod(x, 0j)[1:]
    got = device_fm_demod(x, coef)
    both_nan = np.isnan(want) & np.isnan(got)
    assert not (np.isnan(want) ^ np.isnan(got)).any(), "NaN in one of the two only"
    d = np.abs(want.astype(np.float64) - got)
    d[both_nan] = 0
    assert not ((np.signbit(want) != np.signbit(got)) & ~both_nan).any(), "sign (of zero or of pi) differs"
    return float(d.max())


def test_polynomial_is_within_a_tenth_of_an_ulp_of_pi():
    coef = [np.float64(c) for c in coefficients()]
    t = np.linspace(0, 1, 200001)
    p = np.zeros_like(t) + coef[0]
    for c in coef[1:]:
        p = p * t * t + c
    assert np.abs(p * t - np.arctan(t)).max() < 6e-8


def test_model_against_oracle(port):
    coef = coefficients()
    rng = np.random.default_rng(1)
    n = 400_000
    worst = 0.0
    for sc in (1, 1e-3, 1e3, 1e-10, 3e-19, 1e15):
        x = ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * sc).astype(np.complex64)
        worst = max(worst, compare(port, x, coef))
    b = ((rng.integers(0, 256, 2 * n).astype(f32) - 128) / 128).astype(f32)
    worst = max(worst, compare(port, b.view(np.complex64).copy(), coef))
    b = ((rng.integers(126, 131, 2 * n).astype(f32) - 128) / 128).astype(f32)     # many exact zeros and axis hits
    worst = max(worst, compare(port, b.view(np.complex64).copy(), coef))
    ev = np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, 1e-42, -1e-42, 3e38, -3e38, 1e-30, -1e25], f32)
    soup = rng.choice(ev, 2 * 100_000).astype(f32).view(np.complex64).copy()
    worst = max(worst, compare(port, soup, coef))
    assert worst <= 1e-6, worst
