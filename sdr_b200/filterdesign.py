"""Tap design used by the benchmark and the examples: the Hamming-windowed sinc of ``SDR.FilterDesign``
(hs_sources/SDR/FilterDesign.hs:33-68 -- `sinc`, `hamming`, `windowedSinc`), generalised to even lengths by centring
the window at (n-1)/2 (the reference's `sinc` demands an odd length, FilterDesign.hs:30).  Design-time host arithmetic
in float64, rounded to float32; the result is symmetric, so the Sym constructors can share it.  SDR.FilterDesign itself
stays as it is in the reference (north_star: "Pipes glue and SDR.FilterDesign stay")."""
import numpy as np


def windowed_sinc_taps(n, cutoff, gain=1.0):
    """n taps, cutoff as a fraction of the sample rate (0.5 = Nyquist), DC gain `gain`"""
    k = np.arange(n, dtype=np.float64) - (n - 1) / 2.0
    h = np.sinc(2 * cutoff * k) * 2 * cutoff
    w = 0.54 - 0.46 * np.cos(2 * np.pi * np.arange(n) / (n - 1))
    return (gain * h * w).astype(np.float32)
