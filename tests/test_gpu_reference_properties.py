"""The reference's own QuickCheck properties (reference tests/TestSuite.hs:55-276), re-run with the CUDA implementations
in the set of implementations that must agree.  Generators mirror the reference: buffer sizes 1024..65536, half-tap
counts {32..512} with the full tap list `coeffs ++ reverse coeffs` so symmetric variants can join (:56,:70), factors
from {1,2,3,5,7,11,13,17,23} (:57), values uniform in (-10, 10) (:62).  Agreement is checked at the reference's own
tolerance (absolute 0.01, :284-289) AND at this repo's bar (1e-5 of output scale)."""
import ctypes as C

import numpy as np
import pytest

import oracle.pipes as op

pytestmark = pytest.mark.gpu

SIZES = [1024, 2048, 4096, 8192, 16384, 32768, 65536]
HALF_TAPS = [32, 64, 128, 256, 512]
FACTORS = [1, 2, 3, 5, 7, 11, 13, 17, 23]


@pytest.fixture(scope="module")
def L():
    import sdr_b200
    assert sdr_b200.has_cuda()
    from sdr_b200 import _lib
    return _lib


def _agree(results, ref_tol=0.01):
    """sameResultM (TestSuite.hs:21-26): every implementation agrees with the first"""
    first = results[0][1]
    scale = np.maximum(np.abs(first), np.sqrt(np.mean(np.abs(first) ** 2)))
    for name, r in results[1:]:
        assert r.shape == first.shape, name
        assert np.abs(r - first).max() < ref_tol, f"{name}: {np.abs(r - first).max()} (reference tolerance)"
    for name, r in results:
        if name.startswith("cuda"):
            assert np.all(np.abs(r - first) <= 1e-5 * scale), f"{name}: {float((np.abs(r - first) / scale).max()):.2e} of scale"


def _draw(seed):
    rng = np.random.default_rng(seed)
    size = int(rng.choice(SIZES))
    half = int(rng.choice(HALF_TAPS))
    factor = int(rng.choice(FACTORS))
    coeffs_half = rng.uniform(-10, 10, half).astype(np.float32)
    coeffs = np.concatenate([coeffs_half, coeffs_half[::-1]])
    return rng, size, half, factor, coeffs_half, coeffs


def _call(L, name, num, factor, coeffs, x, cplx):
    xf = L.as_floats(x)
    out = np.zeros(num * (2 if cplx else 1), np.float32)
    c = np.ascontiguousarray(coeffs, np.float32)
    fn = getattr(L.lib, name)
    if name.startswith("filter"):
        L.check(fn(num, len(c), L.ptr(c), L.ptr(xf), L.ptr(out)))
    else:
        L.check(fn(num, factor, len(c), L.ptr(c), L.ptr(xf), L.ptr(out)))
    return out.view(np.complex64) if cplx else out


@pytest.mark.parametrize("seed", range(6))
def test_prop_filters_real(L, ref, seed):
    """propFiltersReal (TestSuite.hs:60-84)"""
    rng, size, half, _, ch, c = _draw(seed)
    x = rng.uniform(-10, 10, size).astype(np.float32)
    num = size - len(c) + 1
    _agree([("filterRR", ref.filter("filterRR", num, c, x)), ("filterSSERR", ref.filter("filterSSERR", num, c, x)),
            ("filterAVXRR", ref.filter("filterAVXRR", num, c, x)), ("filterAVXSymmetricRR", ref.filter("filterAVXSymmetricRR", num, ch, x)),
            ("cudaRR", _call(L, "filterCudaRR", num, 1, c, x, False)),
            ("cudaSymmetricRR", _call(L, "filterCudaSymmetricRR", num, 1, ch, x, False))], ref_tol=0.05)


@pytest.mark.parametrize("seed", range(6))
def test_prop_filters_complex(L, ref, seed):
    """propFiltersComplex (TestSuite.hs:86-111)"""
    rng, size, half, _, ch, c = _draw(100 + seed)
    x = (rng.uniform(-10, 10, size) + 1j * rng.uniform(-10, 10, size)).astype(np.complex64)
    num = size - len(c) + 1
    _agree([("filterRC", ref.filter("filterRC", num, c, x)), ("filterAVXRC", ref.filter("filterAVXRC", num, np.repeat(c, 2), x)),
            ("filterAVXRC2", ref.filter("filterAVXRC2", num, c, x)), ("filterAVXSymmetricRC", ref.filter("filterAVXSymmetricRC", num, ch, x)),
            ("cudaRC", _call(L, "filterCudaRC", num, 1, c, x, True)), ("cudaRCDup", _call(L, "filterCudaRCDup", num, 1, np.repeat(c, 2), x, True)),
            ("cudaSymmetricRC", _call(L, "filterCudaSymmetricRC", num, 1, ch, x, True))], ref_tol=0.05)


@pytest.mark.parametrize("seed", range(8))
def test_prop_decimators(L, ref, seed):
    """propDecimationReal / propDecimationComplex (TestSuite.hs:113-165)"""
    rng, size, half, factor, ch, c = _draw(200 + seed)
    num = (size - len(c) + 1) // factor
    if num <= 0:
        pytest.skip("draw leaves no output")
    x = rng.uniform(-10, 10, size).astype(np.float32)
    _agree([("decimateRR", ref.decimate("decimateRR", num, factor, c, x)), ("decimateAVXRR", ref.decimate("decimateAVXRR", num, factor, c, x)),
            ("decimateAVXSymmetricRR", ref.decimate("decimateAVXSymmetricRR", num, factor, ch, x)),
            ("cudaRR", _call(L, "decimateCudaRR", num, factor, c, x, False)),
            ("cudaSymmetricRR", _call(L, "decimateCudaSymmetricRR", num, factor, ch, x, False))], ref_tol=0.05)
    z = (rng.uniform(-10, 10, size) + 1j * rng.uniform(-10, 10, size)).astype(np.complex64)
    _agree([("decimateRC", ref.decimate("decimateRC", num, factor, c, z)), ("decimateAVXRC", ref.decimate("decimateAVXRC", num, factor, np.repeat(c, 2), z)),
            ("decimateAVXSymmetricRC", ref.decimate("decimateAVXSymmetricRC", num, factor, ch, z)),
            ("cudaRC", _call(L, "decimateCudaRC", num, factor, c, z, True)),
            ("cudaRCDup", _call(L, "decimateCudaRCDup", num, factor, np.repeat(c, 2), z, True)),
            ("cudaSymmetricRC", _call(L, "decimateCudaSymmetricRC", num, factor, ch, z, True))], ref_tol=0.05)


@pytest.mark.parametrize("seed", range(8))
def test_prop_resamplers(L, ref, seed):
    """propResamplingReal / Complex incl. a random start group (TestSuite.hs:167-227); decimation > interpolation (:173)"""
    rng, size, half, _, ch, c = _draw(300 + seed)
    decim = int(rng.choice([f for f in FACTORS if f > 1]))
    interp = int(rng.choice([f for f in FACTORS if f < decim]))
    num_coeffs, increments, groups = op.prepare_coeffs(8, interp, decim, c)
    start = int(rng.integers(0, len(increments)))
    num = (size * interp - op.round_up(len(c), interp * 8)) // decim - len(increments) - 1
    if num <= 0:
        pytest.skip("draw leaves no output")
    inc = np.ascontiguousarray(increments, np.int32)
    rows = (C.POINTER(C.c_float) * groups.shape[0])(*[groups[i].ctypes.data_as(C.POINTER(C.c_float)) for i in range(groups.shape[0])])
    for cplx in (False, True):
        x = rng.uniform(-10, 10, size).astype(np.float32)
        if cplx:
            x = (x + 1j * rng.uniform(-10, 10, size)).astype(np.complex64)
        names = ("resample2RC", "resampleAVXRC") if cplx else ("resample2RR", "resampleAVXRR")
        res = [(n, ref.resample(n, num, num_coeffs, start, increments, groups, x)) for n in names]
        out = np.zeros(num * (2 if cplx else 1), np.float32)
        xf = L.as_floats(x)
        g = L.check_group((L.lib.resampleCudaRC if cplx else L.lib.resampleCudaRR)(num, num_coeffs, start, len(increments), L.ptr(inc), rows,
                                                                                   L.ptr(xf), L.ptr(out)))
        assert all(g == r[1] for _, r in res)
        _agree([(n, r[0]) for n, r in res] + [("cuda", out.view(np.complex64) if cplx else out)], ref_tol=0.05)


@pytest.mark.parametrize("seed", range(3))
def test_prop_conversions_and_scaling(L, ref, seed):
    """propConversion / propConversionBladeRF / propScaleReal (TestSuite.hs:229-276)"""
    import sdr_b200
    rng = np.random.default_rng(400 + seed)
    size = int(rng.choice(SIZES))
    b = rng.integers(0, 256, 2 * size, dtype=np.uint8)
    assert np.array_equal(sdr_b200.interleavedIQUnsignedByteToFloat(b).view(np.float32), ref.convert_u8("convertCAVX", b))
    assert np.array_equal(ref.convert_u8("convertC", b), ref.convert_u8("convertCSSE", b))
    v = rng.integers(-2048, 2048, 2 * size, dtype=np.int16)
    assert np.array_equal(sdr_b200.interleavedIQSigned2048ToFloat(v).view(np.float32), ref.convert_i16("convertCAVXBladeRF", v))
    x = rng.uniform(-10, 10, size).astype(np.float32)
    k = np.float32(rng.uniform(-10, 10))
    assert np.array_equal(sdr_b200.scaleFast(k, x), ref.scale("scaleAVX", k, x))


def test_edge_cases(L):
    """empty calls, single outputs, minimum-length inputs, invalid arguments"""
    import sdr_b200
    c = np.ones(8, np.float32)
    x = np.arange(64, dtype=np.float32)
    out = np.zeros(1, np.float32)
    L.check(L.lib.filterCudaRR(0, 8, L.ptr(c), L.ptr(x), L.ptr(out)))           # num = 0: no-op
    L.check(L.lib.filterCudaRR(1, 8, L.ptr(c), L.ptr(x), L.ptr(out)))
    assert out[0] == np.float32(sum(range(8)))
    assert L.lib.filterCudaRR(-1, 8, L.ptr(c), L.ptr(x), L.ptr(out)) == L.SDR_EINVAL
    assert L.lib.decimateCudaRR(4, 0, 8, L.ptr(c), L.ptr(x), L.ptr(out)) == L.SDR_EINVAL
    assert L.lib.decimateCudaRCDup(1, 2, 7, L.ptr(c), L.ptr(x), L.ptr(out)) == L.SDR_EINVAL   # odd duplicated length
    assert b"decimateCudaRCDup" in L.lib.sdr_last_error()
    assert len(sdr_b200.interleavedIQUnsignedByteToFloat(np.zeros(0, np.uint8))) == 0
    assert len(sdr_b200.scaleFast(2.0, np.zeros(0, np.float32))) == 0
    # a decimator whose window is the whole (minimum-length) vector: exactly one output
    d = sdr_b200.cudaDecimatorR(5, c)
    y = d.decimateOne(1, x[:8])
    assert y[0] == np.float32(sum(range(8)))
    # tap counts that are not a multiple of any SIMD width, prime decimation
    taps = np.linspace(-1, 1, 37).astype(np.float32)
    z = (np.arange(1000) % 17 - 8).astype(np.float32)
    d = sdr_b200.cudaDecimatorR(23, taps)
    got = d.decimateOne(41, z)
    want = op.flat_decimate(z, taps, 23, 41).astype(np.float32)
    assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max()
