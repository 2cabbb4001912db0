"""CPU replica of the library's counter-based synthetic streams (sdr_synth_noise / sdr_synth_bytes,
sdr_b200/csrc/kernels_generic.cu) and of its position-weighted checksum -- test infrastructure."""
import numpy as np


def _mix32(x):
    x = x.astype(np.uint32)
    x ^= x >> np.uint32(16)
    x = (x * np.uint32(0x7feb352d)).astype(np.uint32)
    x ^= x >> np.uint32(15)
    x = (x * np.uint32(0x846ca68b)).astype(np.uint32)
    x ^= x >> np.uint32(16)
    return x


def noise(n, first=0, seed=0x5D2B200):
    idx = np.arange(first, first + n, dtype=np.uint64)
    lo = (idx & np.uint64(0xffffffff)).astype(np.uint32)
    hi = (idx >> np.uint64(32)).astype(np.uint32)
    with np.errstate(over="ignore"):
        h1 = _mix32(lo ^ _mix32(hi ^ np.uint32(seed)))
        h2 = _mix32(h1 ^ np.uint32(0x9E3779B9))
    s = ((h1 & np.uint32(0xffff)).astype(np.int64) + (h1 >> np.uint32(16)).astype(np.int64) +
         (h2 & np.uint32(0xffff)).astype(np.int64) + (h2 >> np.uint32(16)).astype(np.int64) - 131070)
    return (s.astype(np.float32) * np.float32(2.6429e-5)).astype(np.float32)


def noise_complex(n, first=0, seed=0x5D2B200):
    return noise(2 * n, 2 * first, seed).view(np.complex64)


def rand_bytes(n, first=0, seed=0x5D2B200):
    g = np.arange(first, first + n, dtype=np.uint64)
    w = g >> np.uint64(2)
    lo = (w & np.uint64(0xffffffff)).astype(np.uint32)
    hi = (w >> np.uint64(32)).astype(np.uint32)
    with np.errstate(over="ignore"):
        h = _mix32(lo ^ _mix32(hi ^ np.uint32(seed) ^ np.uint32(0xB5297A4D)))
    return ((h >> (np.uint32(8) * (g & np.uint64(3)).astype(np.uint32))) & np.uint32(0xff)).astype(np.uint8)


def checksum32(words, first=0):
    w = np.ascontiguousarray(words).view(np.uint32).astype(np.uint64)
    idx = np.arange(first, first + len(w), dtype=np.uint64)
    with np.errstate(over="ignore"):
        return int(np.sum((w + np.uint64(1)) * (np.uint64(2) * idx + np.uint64(1)), dtype=np.uint64))


def windowed_sinc_taps(n, cutoff, gain=1.0):
    """the package's tap designer (sdr_b200/filterdesign.py), loaded by path so that CPU-only users of this module do not
    need the native library"""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location(
        "_sdr_b200_filterdesign", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "sdr_b200", "filterdesign.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.windowed_sinc_taps(n, cutoff, gain)
