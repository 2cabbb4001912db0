"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI, against the oracle
(oracle/liboracle.so = our C restatement, oracle/_ref/libsdrref.so = the unmodified reference C) on the same seeded
inputs.

Tolerances.  FAST arithmetic (fused multiply-add, taps in increasing order): BASELINE.json's "1e-5 relative", read as
SURVEY.md section 8c does: |y - y_ref| <= 1e-5 * max(|y_ref|, rms(y_ref)).  EXACT arithmetic, the u8 / i16 converts, scale
and the dc blocker: bit-exact."""
import ctypes as C

import numpy as np
import pytest

import oracle
import oracle.pipes as op
import synth
from oracle import V_AVX, V_AVX2, V_AVXSYM, V_SCALAR, V_SSE, V_SSE2, V_SSESYM

pytestmark = pytest.mark.gpu

RTOL = 1e-5


def close(y, ref, rtol=RTOL):
    y, ref = np.asarray(y), np.asarray(ref)
    assert y.shape == ref.shape, (y.shape, ref.shape)
    if y.size == 0:
        return True
    scale = np.maximum(np.abs(ref), np.sqrt(np.mean(np.abs(ref) ** 2)))
    err = np.abs(y - ref)
    bad = err > rtol * scale
    assert not bad.any(), f"max err/scale {float((err / np.maximum(scale, 1e-30)).max()):.3e} at {int(np.argmax(err / np.maximum(scale, 1e-30)))}"
    return True


@pytest.fixture(scope="module")
def sdr():
    import sdr_b200
    assert sdr_b200.has_cuda(), "no sm_100 device: the product has no CPU path"
    return sdr_b200


@pytest.fixture(scope="module")
def L():
    from sdr_b200 import _lib
    return _lib


@pytest.fixture(scope="module")
def ctx(sdr):
    return sdr.default_context()


def rnd(n, cplx, seed):
    rng = np.random.default_rng(seed)
    if cplx:
        return (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    return rng.standard_normal(n).astype(np.float32)


def taps_for(n, seed):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal(n) / np.sqrt(n)).astype(np.float32)


# ---------------------------------------------------------------------------------------------------------------
# layer 1: reference-signature one-shot entry points, FAST arithmetic vs the reference AVX C
# ---------------------------------------------------------------------------------------------------------------
ONE_SHOT = [
    # (cuda symbol, ref symbol, complex, coefficient form)
    ("filterCudaRR", "filterAVXRR", False, "plain"),
    ("filterCudaSymmetricRR", "filterAVXSymmetricRR", False, "sym"),
    ("filterCudaRC", "filterAVXRC2", True, "plain"),
    ("filterCudaRCDup", "filterAVXRC", True, "dup"),
    ("filterCudaSymmetricRC", "filterAVXSymmetricRC", True, "sym"),
    ("decimateCudaRR", "decimateAVXRR", False, "plain"),
    ("decimateCudaSymmetricRR", "decimateAVXSymmetricRR", False, "sym"),
    ("decimateCudaRC", "decimateAVXRC2", True, "plain"),
    ("decimateCudaRCDup", "decimateAVXRC", True, "dup"),
    ("decimateCudaSymmetricRC", "decimateAVXSymmetricRC", True, "sym"),
]


def _coeff_forms(T, seed):
    half = taps_for(T // 2, seed)
    full = np.concatenate([half, half[::-1]])
    return {"plain": full, "dup": np.repeat(full, 2), "sym": half}, full


@pytest.mark.parametrize("cuda_name,ref_name,cplx,form", ONE_SHOT)
@pytest.mark.parametrize("T,factor,n_in", [(128, 8, 8192), (64, 1, 8192), (16, 3, 1000), (256, 5, 4096 + 13)])
def test_oneshot_fast_vs_reference(L, ref, port, cuda_name, ref_name, cplx, form, T, factor, n_in):
    if cuda_name.startswith("filter"):
        factor = 1
    forms, full = _coeff_forms(T, seed=T + factor)
    x = rnd(n_in, cplx, seed=n_in + T)
    num = (n_in - T) // factor + 1
    coeffs = forms[form]
    xf = L.as_floats(x)
    out = np.zeros(num * (2 if cplx else 1), np.float32)
    fn = getattr(L.lib, cuda_name)
    if cuda_name.startswith("filter"):
        L.check(fn(num, len(coeffs), L.ptr(coeffs), L.ptr(xf), L.ptr(out)))
        want = ref.filter(ref_name, num, coeffs, x)
    else:
        L.check(fn(num, factor, len(coeffs), L.ptr(coeffs), L.ptr(xf), L.ptr(out)))
        want = ref.decimate(ref_name, num, factor, coeffs, x)
    got = out.view(np.complex64) if cplx else out
    close(got, want)
    # and against the float64 flat-stream model
    close(got, op.flat_decimate(x, full, factor, num).astype(got.dtype), rtol=2e-6)


def test_oneshot_headline_block_uses_tuned_kernel(sdr, ref):
    """cfg2: 8192-sample complex block, 128 taps, decimate by 8 -> 1009 outputs in the C call (Filter.hs:588)"""
    taps = synth.windowed_sinc_taps(128, 1 / 16)
    x = synth.noise_complex(8192)
    d = sdr.cudaDecimatorC(8, taps, sizeMultiple=4)
    assert d.numCoeffsD == 128 and d.decimationD == 8
    got = d.decimateOne(1009, x)
    assert d.last_kernel().startswith("dec_c_ring")
    want = ref.decimate("decimateAVXRC", 1009, 8, np.repeat(taps, 2), x)
    close(got, want)


# ---------------------------------------------------------------------------------------------------------------
# EXACT arithmetic: bit-identical to every reference variant
# ---------------------------------------------------------------------------------------------------------------
EXACT_R = [(V_SCALAR, "decimateRR", "plain"), (V_SSE, "decimateSSERR", "plain"), (V_AVX, "decimateAVXRR", "plain"),
           (V_SSESYM, "decimateSSESymmetricRR", "sym"), (V_AVXSYM, "decimateAVXSymmetricRR", "sym")]
EXACT_C = [(V_SCALAR, "decimateRC", "plain"), (V_SSE, "decimateSSERC", "dup"), (V_AVX, "decimateAVXRC", "dup"),
           (V_SSE2, "decimateSSERC2", "plain"), (V_AVX2, "decimateAVXRC2", "plain"),
           (V_SSESYM, "decimateSSESymmetricRC", "sym"), (V_AVXSYM, "decimateAVXSymmetricRC", "sym")]


@pytest.mark.parametrize("cplx,variant,ref_name,form", [(False,) + e for e in EXACT_R] + [(True,) + e for e in EXACT_C])
@pytest.mark.parametrize("T,factor,n_in", [(128, 8, 8192), (64, 1, 2048), (32, 7, 3000)])
def test_exact_bitwise_vs_reference(L, ref, port, cplx, variant, ref_name, form, T, factor, n_in):
    forms, _ = _coeff_forms(T, seed=7 * T + factor)
    coeffs = forms[form]
    x = rnd(n_in, cplx, seed=n_in * 3 + T)
    num = (n_in - T) // factor + 1
    xf = L.as_floats(x)
    out = np.zeros(num * (2 if cplx else 1), np.float32)
    L.check(L.lib.sdr_exact_decimate(variant, int(cplx), num, factor, len(coeffs), L.ptr(coeffs), L.ptr(xf), L.ptr(out)))
    want = ref.decimate(ref_name, num, factor, coeffs, x)
    got = out.view(np.complex64) if cplx else out
    assert np.array_equal(got, want), f"{ref_name}: {int((got != want).sum())} of {num} differ"
    assert np.array_equal(got, port.decimate(variant, num, factor, coeffs, x, cplx))


@pytest.mark.parametrize("cplx,variant,ref_name", [(False, V_SCALAR, "resample2RR"), (False, V_SSE, "resampleSSERR"),
                                                   (False, V_AVX, "resampleAVXRR"), (True, V_SCALAR, "resample2RC"),
                                                   (True, V_SSE2, "resampleSSERC"), (True, V_AVX2, "resampleAVXRC")])
@pytest.mark.parametrize("interp,decim,T,start", [(3, 10, 90, 0), (3, 10, 90, 2), (2, 7, 64, 1), (5, 11, 128, 3)])
def test_exact_resample_bitwise_vs_reference(L, ref, cplx, variant, ref_name, interp, decim, T, start):
    sm = {V_SCALAR: 1, V_SSE: 4, V_AVX: 8, V_SSE2: 4, V_AVX2: 8}[variant]
    taps = taps_for(T, seed=T + interp)
    num_coeffs, increments, groups = op.prepare_coeffs(sm, interp, decim, taps)
    n_in = 4096
    x = rnd(n_in, cplx, seed=decim)
    num = (n_in * interp - op.round_up(T, interp * sm)) // decim - 8
    want, g_want = ref.resample(ref_name, num, num_coeffs, start, increments, groups, x)
    xf = L.as_floats(x)
    out = np.zeros(num * (2 if cplx else 1), np.float32)
    inc = np.ascontiguousarray(increments, np.int32)
    g = C.c_int()
    # the reference's SIMD loops run over the zero padding: hand the padded row length as num_coeffs
    L.check(L.lib.sdr_exact_resample(variant, int(cplx), num, groups.shape[1], start, len(increments), L.ptr(inc),
                                     L.ptr(groups), groups.shape[1], L.ptr(xf), L.ptr(out), C.byref(g)))
    got = out.view(np.complex64) if cplx else out
    assert g.value == g_want
    assert np.array_equal(got, want), f"{ref_name}: {int((got != want).sum())} of {num} differ"


# ---------------------------------------------------------------------------------------------------------------
# resampler FAST, converts, scale, fm demod, dc blocker
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("interp,decim,T", [(3, 10, 90), (2, 7, 64), (7, 11, 100)])
def test_resample_fast_vs_reference(L, ref, cplx, interp, decim, T):
    taps = taps_for(T, seed=T)
    num_coeffs, increments, groups = op.prepare_coeffs(8, interp, decim, taps)
    x = rnd(8192, cplx, seed=interp * decim)
    num = (8192 * interp - op.round_up(T, interp * 8)) // decim + 1
    want, g_want = ref.resample("resampleAVXRC" if cplx else "resampleAVXRR", num, num_coeffs, 0, increments, groups, x)
    rows = (C.POINTER(C.c_float) * groups.shape[0])(*[groups[i].ctypes.data_as(C.POINTER(C.c_float)) for i in range(groups.shape[0])])
    xf = L.as_floats(x)
    out = np.zeros(num * (2 if cplx else 1), np.float32)
    inc = np.ascontiguousarray(increments, np.int32)
    fn = L.lib.resampleCudaRC if cplx else L.lib.resampleCudaRR
    # the reference's own signature: 8 arguments, returns the next group (resample.c:70-87)
    g = L.check_group(fn(num, num_coeffs, 0, len(increments), L.ptr(inc), rows, L.ptr(xf), L.ptr(out)))
    got = out.view(np.complex64) if cplx else out
    assert g == g_want
    assert fn(num, num_coeffs, len(increments), len(increments), L.ptr(inc), rows, L.ptr(xf), L.ptr(out)) == -L.SDR_EINVAL
    close(got, want)
    close(got, op.flat_resample(x, taps, interp, decim, num).astype(got.dtype), rtol=2e-6)


def test_resample_legacy_vs_reference(L, ref):
    taps = taps_for(90, 5)
    x = rnd(4096, False, 9)
    for off in (0, 1, 2):
        num = 1200
        want = ref.resample_legacy(num, 3, 10, off, taps, x)
        out = np.zeros(num, np.float32)
        L.check(L.lib.resampleCudaLegacyRR(num, len(taps), 3, 10, off, L.ptr(taps), L.ptr(x), L.ptr(out)))
        close(out, want)


@pytest.mark.parametrize("n", [0, 2, 30, 16384, 16384 + 6, 1 << 20])
def test_convert_u8_bitexact(sdr, ref, n):
    b = np.random.default_rng(n).integers(0, 256, n, dtype=np.uint8)
    got = sdr.interleavedIQUnsignedByteToFloat(b)
    # the reference's SIMD converts step 8 / 16 elements with no tail handling (convert.c:37-50): only the scalar
    # one is defined for ragged sizes
    want = ref.convert_u8("convertCAVX" if n % 16 == 0 else "convertC", b) if n else np.zeros(0, np.float32)
    assert np.array_equal(got.view(np.float32), want)
    if n == 30:   # every byte value
        allb = np.arange(256, dtype=np.uint8)
        assert np.array_equal(sdr.interleavedIQUnsignedByteToFloat(allb).view(np.float32), ref.convert_u8("convertC", allb))


def test_convert_bladerf_bitexact(sdr, ref):
    v = np.random.default_rng(1).integers(-2048, 2048, 8192, dtype=np.int16)
    assert np.array_equal(sdr.interleavedIQSigned2048ToFloat(v).view(np.float32), ref.convert_i16("convertCAVXBladeRF", v))
    x = np.random.default_rng(2).uniform(-1.2, 1.2, 8192).astype(np.float32)
    got = sdr.complexFloatToInterleavedIQSigned2048(x.view(np.complex64))
    assert np.array_equal(got, ref.convert_tx(x))


@pytest.mark.parametrize("n", [1, 7, 8192, 8192 + 3])
def test_scale_bitexact(sdr, ref, n):
    x = rnd(n, False, n)
    # scaleAVX steps 8 floats with no tail handling (scale.c:30-36): scalar reference for ragged sizes
    assert np.array_equal(sdr.scaleFast(0.2, x), ref.scale("scaleAVX" if n % 8 == 0 else "scale", np.float32(0.2), x))


def test_fm_demod_vs_oracle(sdr, port):
    x = rnd(8192, True, 3)
    x[100] = 0  # phase 0 = 0 branch
    x[200] = x[199] * np.complex64(-2.0)   # phase pi
    got = sdr.fmDemodVec(0j, x)
    want = port.fm_demod(x, 0j)
    assert np.abs(got - want).max() <= 1e-5   # |phase| <= pi: 1e-5 relative of the output range
    # carried last sample
    got2 = sdr.fmDemodVec(x[-1], x[:50])
    assert np.abs(got2 - port.fm_demod(x[:50], x[-1])).max() <= 1e-5
    # streaming form == one-shot on the flat stream
    outs = list(sdr.fmDemod([x[:1000], x[1000:1001], x[1001:]]))
    assert np.array_equal(np.concatenate(outs), got)


def test_dc_blocking_filter_pipe_bitexact(sdr, ref):
    """dcBlockingFilter (Filter.hs:730-739): state carried across vectors; bit-exact vs the reference C run on the flat stream"""
    x = rnd(3 * 4096 + 77, False, 12)
    chunks = [x[:4096], x[4096:4096 + 77], x[4096 + 77:]]
    got = np.concatenate(list(sdr.dcBlockingFilter(chunks)))
    want, _, _ = ref.dc_blocker(x, 0.0, 0.0)
    assert np.array_equal(got, want)


def test_dc_blocker_bitexact(sdr, ref):
    x = rnd(8192 + 5, False, 11)
    got, fs, fo = sdr.dcBlocker(x, 0.25, -0.5)
    want, ws, wo = ref.dc_blocker(x, 0.25, -0.5)
    assert np.array_equal(got, want) and fs == ws and fo == wo


def _dc_dev(ctx, x, s0=0.0, o0=0.0, in_off=0, out_off=0):
    """sdr_dev_dc_blocker on device buffers placed in_off / out_off floats past a 256-byte aligned allocation"""
    n = len(x)
    d_in, d_out, d_fin = ctx.alloc(4 * n + 64), ctx.alloc(4 * n + 64), ctx.alloc(8)
    d_in.upload(x, 4 * in_off)
    ctx.dc_blocker(d_in.at(4 * in_off), d_out.at(4 * out_off), n, d_fin.ptr, s0, o0)
    got, fin = d_out.to_host(np.float32, n, 4 * out_off), d_fin.to_host(np.float32, 2)
    for b in (d_in, d_out, d_fin):
        b.free()
    return got, fin


def _dc_delta(before, after):
    return tuple(a - b for a, b in zip(after, before))


def test_dc_blocker_parallel_bitexact(sdr, ctx, ref):
    """long vectors take the speculative chunk-parallel path (csrc/dc_spec.cuh): bit-exact, and with the default warm-up
    no chunk of a noise stream needs a repair"""
    x = rnd(1_000_003, False, 21)
    want, ws, wo = ref.dc_blocker(x, 0.25, -0.5)
    st0, _ = ctx.dc_stats()
    got, fin = _dc_dev(ctx, x, 0.25, -0.5)
    st1, par = ctx.dc_stats()
    assert par, "a 1M-sample vector must take the parallel path"
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert np.array_equal(fin.view(np.uint32), np.array([ws, wo], np.float32).view(np.uint32))
    calls, chunks, repaired, _ = _dc_delta(st0, st1)
    assert calls == 1 and chunks >= 60 and repaired == 0, (calls, chunks, repaired)
    # the host-pointer entry point of layer 1 goes the same way
    got2, fs, fo = sdr.dcBlocker(x, 0.25, -0.5)
    assert np.array_equal(got2.view(np.uint32), want.view(np.uint32)) and fs == ws and fo == wo
    # 4-byte aligned device pointers: scalar-access instantiation
    got3, fin3 = _dc_dev(ctx, x[:300_001], 0.25, -0.5, in_off=1, out_off=3)
    want3, ws3, wo3 = ref.dc_blocker(x[:300_001], 0.25, -0.5)
    assert ctx.dc_stats()[1]
    assert np.array_equal(got3.view(np.uint32), want3.view(np.uint32)) and fin3[1] == wo3 and fin3[0] == ws3


def test_dc_blocker_parallel_repairs(ctx, ref):
    """speculation that misses (warm-up far too short; constant input parked on a denormal fixed point) is repaired
    serially: still bit-exact"""
    x = (3.0 * rnd(400_000, False, 22) + 1.0).astype(np.float32)
    want, ws, wo = ref.dc_blocker(x, 0.0, 7.0)
    try:
        for chunk, k1, k2, lo in ((4096, 0, 32, 80), (512, 0, 0, 700), (2048, 6144, 512, 1)):
            ctx.dc_tuning(chunk, k1, k2, 0)
            st0, _ = ctx.dc_stats()
            got, fin = _dc_dev(ctx, x, 0.0, 7.0)
            st1, par = ctx.dc_stats()
            assert par
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (chunk, k1, k2)
            assert fin[0] == ws and fin[1] == wo
            assert _dc_delta(st0, st1)[2] >= lo, (chunk, k1, k2, _dc_delta(st0, st1))
        ctx.dc_tuning(2048, 1024, 1024, 0)
        c = np.full(150_000, 0.5, np.float32)
        wantc, _, woc = ref.dc_blocker(c, 0.0, 0.0)
        gotc, finc = _dc_dev(ctx, c)
        assert np.array_equal(gotc.view(np.uint32), wantc.view(np.uint32)) and finc[1] == woc
        assert 0 < abs(float(woc)) < 1e-42
        # ragged lengths through the parallel path (min_parallel = 0)
        ctx.dc_tuning(64, 64, 512, 0)
        for n in (1, 7, 9, 1025, 4099):
            w, _, wo_n = ref.dc_blocker(x[:n], 0.1, 0.2)
            g, f = _dc_dev(ctx, x[:n], 0.1, 0.2)
            assert np.array_equal(g.view(np.uint32), w.view(np.uint32)) and f[1] == wo_n, n
    finally:
        ctx.dc_tuning()


def test_dc_blocking_filter_pipe_long_vectors(sdr, ctx, ref):
    """dcBlockingFilter with vectors long enough for the parallel path: (lastSample, lastOutput) carried on the device
    from one vector to the next"""
    x = rnd(3 * 131072 + 70_000, False, 23)
    cuts = [0, 131072, 131072 + 70_000, 131072 + 70_000 + 100, len(x)]
    chunks = [x[a:b] for a, b in zip(cuts[:-1], cuts[1:])]
    got = np.concatenate(list(sdr.dcBlockingFilter(chunks, ctx)))
    want, _, _ = ref.dc_blocker(x, 0.0, 0.0)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_dc_blocker_full_size_parallel_equals_serial(ctx):
    """size-independent property at 2^25 samples: the chunk-parallel evaluation and the serial kernel (forced by
    min_parallel) write the same words (checksum of the device buffers)"""
    n = 1 << 25
    d_in, d_out, d_fin = ctx.alloc(4 * n), ctx.alloc(4 * n), ctx.alloc(8)
    ctx.synth_noise(d_in, n)
    try:
        ctx.dc_blocker(d_in.ptr, d_out.ptr, n, d_fin.ptr)
        assert ctx.dc_stats()[1]
        par = (ctx.checksum32(d_out, n), d_fin.to_host(np.uint32, 2).tolist())
        ctx.dc_tuning(0, -1, -1, 1 << 40)
        ctx.dc_blocker(d_in.ptr, d_out.ptr, n, d_fin.ptr)
        assert not ctx.dc_stats()[1]
        ser = (ctx.checksum32(d_out, n), d_fin.to_host(np.uint32, 2).tolist())
        assert par == ser
    finally:
        ctx.dc_tuning()
        for b in (d_in, d_out, d_fin):
            b.free()


# ---------------------------------------------------------------------------------------------------------------
# layer 2: record closures under the reference's own state machine; layer 3: native pipes vs the flat stream
# ---------------------------------------------------------------------------------------------------------------
def _chunks(x, sizes):
    out, i = [], 0
    for s in sizes:
        out.append(x[i:i + s])
        i += s
    assert i == len(x)
    return out


@pytest.mark.parametrize("cplx", [False, True])
def test_record_closures_in_reference_state_machine(sdr, cplx):
    """our filterOne/filterCross etc. dropped into the restated firDecimator/firFilter state machines
    (oracle/pipes.py, Filter.hs:536-611) give the same stream as the oracle's own AVX records"""
    taps = taps_for(128, 1)
    x = rnd(8192 * 4, cplx, 2)
    sizes = [8192, 5000, 11192, 8384]
    mk_ours = (sdr.cudaDecimatorC if cplx else sdr.cudaDecimatorR)
    ours = mk_ours(8, taps, sizeMultiple=4 if cplx else 8)
    theirs = (op.mk_decimator_c if cplx else op.mk_decimator)(V_AVX, 8, taps)
    assert ours.numCoeffsD == theirs.numCoeffsD
    a = np.concatenate(list(op.fir_decimator(op.Decimator(ours.numCoeffsD, 8, ours.decimateOne, ours.decimateCross, cplx),
                                             512, _chunks(x, sizes))))
    b = np.concatenate(list(op.fir_decimator(theirs, 512, _chunks(x, sizes))))
    close(a, b)
    fo = (sdr.cudaFilterC if cplx else sdr.cudaFilterR)(taps[:64])
    ft = (op.mk_filter_c if cplx else op.mk_filter)(V_AVX, taps[:64])
    a = np.concatenate(list(op.fir_filter(op.Filter(fo.numCoeffsF, fo.filterOne, fo.filterCross, cplx), 4096, _chunks(x, sizes))))
    b = np.concatenate(list(op.fir_filter(ft, 4096, _chunks(x, sizes))))
    close(a, b)


@pytest.mark.parametrize("cplx", [False, True])
def test_resampler_record_in_reference_state_machine(sdr, cplx):
    taps = taps_for(90, 4)
    x = rnd(8192 * 3, cplx, 5)
    sizes = [8192, 3000, 13384]
    ours = (sdr.cudaResamplerC if cplx else sdr.cudaResamplerR)(3, 10, taps, sizeMultiple=8)
    theirs = op.mk_resampler(V_AVX2 if cplx else V_AVX, 3, 10, taps, cplx=cplx)
    assert ours.numCoeffsR == theirs.numCoeffsR
    mine = op.Resampler(ours.numCoeffsR, 10, 3, (0, 0), ours.resampleOne, ours.resampleCross, cplx)
    a = np.concatenate(list(op.fir_resampler(mine, 1024, _chunks(x, sizes))))
    b = np.concatenate(list(op.fir_resampler(theirs, 1024, _chunks(x, sizes))))
    close(a, b)


@pytest.mark.parametrize("cplx,T,D,block_out,sizes", [
    (True, 128, 8, 8192, [8192] * 16),                       # cfg2: one output vector per 8 input vectors
    (True, 128, 8, 1000, [8192, 3001, 300, 9000, 20000]),    # ragged (the reference itself asserts on vectors too close to numCoeffs)
    (False, 128, 8, 512, [8192, 777, 4096]),
    (True, 51, 8, 256, [4096] * 5),                          # the FM example's RF decimator length
])
def test_native_decimator_pipe_vs_reference_pipe(sdr, cplx, T, D, block_out, sizes):
    taps = taps_for(T, T)
    x = rnd(sum(sizes), cplx, 6)
    sm = 4 if cplx else 8
    d = (sdr.cudaDecimatorC if cplx else sdr.cudaDecimatorR)(D, taps, sizeMultiple=sm)
    got = list(sdr.firDecimator(d, block_out, _chunks(x, sizes)))
    want = list(op.fir_decimator((op.mk_decimator_c if cplx else op.mk_decimator)(V_AVX, D, taps), block_out, _chunks(x, sizes)))
    assert len(got) == len(want) and all(len(g) == block_out for g in got)
    if got:
        close(np.concatenate(got), np.concatenate(want))


@pytest.mark.parametrize("sizes", [[8192] * 4, [8192, 200, 5000, 150, 130, 9999]])
def test_native_filter_pipe_vs_reference_pipe(sdr, sizes):
    half = taps_for(32, 8)
    x = rnd(sum(sizes), False, 7)
    f = sdr.cudaFilterSymR(half)
    assert f.numCoeffsF == 64
    got = list(sdr.firFilter(f, 4096, _chunks(x, sizes)))
    want = list(op.fir_filter(op.mk_filter_sym_r(V_AVXSYM, half), 4096, _chunks(x, sizes)))
    assert len(got) == len(want)
    close(np.concatenate(got), np.concatenate(want))


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("sizes", [[8192] * 4, [8192, 200, 5000, 133, 12000]])
def test_native_resampler_pipe_vs_reference_pipe(sdr, cplx, sizes):
    taps = taps_for(90, 9)
    x = rnd(sum(sizes), cplx, 8)
    r = (sdr.cudaResamplerC if cplx else sdr.cudaResamplerR)(3, 10, taps, sizeMultiple=8)
    got = list(sdr.firResampler(r, 1024, _chunks(x, sizes)))
    want = list(op.fir_resampler(op.mk_resampler(V_AVX2 if cplx else V_AVX, 3, 10, taps, cplx=cplx), 1024, _chunks(x, sizes)))
    assert len(got) == len(want)
    close(np.concatenate(got), np.concatenate(want))


def test_pipe_run_pinned_continuing_stream(sdr, L, ref):
    """the e2e path of bench.py: sdr_pipe_run with SDR_HOST_PINNED vectors in and out, batching knob set, several passes
    through one long-lived pipe (the stream continues across calls); every yielded vector against the reference C"""
    import ctypes as C
    BUF, n_vecs, passes = 8192, 48, 3
    taps = synth.windowed_sinc_taps(128, 1 / 16)
    dec = sdr.cudaDecimatorC(8, taps, sizeMultiple=4)
    pipe = sdr.pipeFirDecimator(dec, BUF)
    L.check(L.lib.sdr_pipe_set_batch(pipe.h, 4 * BUF))
    hin = sdr.PinnedArray(np.float32, 2 * n_vecs * BUF)
    hout = sdr.PinnedArray(np.float32, 2 * (n_vecs * BUF // 8 + BUF))
    stream = synth.noise_complex(passes * n_vecs * BUF)
    got = []
    n_out = C.c_longlong()
    for k in range(passes):
        hin.array[:] = stream[k * n_vecs * BUF:(k + 1) * n_vecs * BUF].view(np.float32)
        L.check(L.lib.sdr_pipe_run(pipe.h, pipe.h, hin.p, BUF, n_vecs, L.SDR_HOST_PINNED, hout.p, len(hout.array) // 2,
                                   L.SDR_HOST_PINNED, C.byref(n_out)))
        assert n_out.value % BUF == 0
        got.append(hout.array[:2 * n_out.value].view(np.complex64).copy())
    y = np.concatenate(got)
    total = (len(stream) - 128) // 8 + 1
    assert len(y) == (total // BUF) * BUF       # only whole vectors are ever yielded
    want = ref.decimate("decimateAVXRC", len(y), 8, np.repeat(taps, 2), stream)
    close(y, want)
    pipe.close(); hin.free(); hout.free()


def test_pipe_precondition_errors(sdr):
    d = sdr.cudaDecimatorC(8, taps_for(128, 1), sizeMultiple=4)
    with pytest.raises(sdr.SdrError) as e:
        list(sdr.firDecimator(d, 1024, [np.zeros(100, np.complex64)]))
    assert e.value.code == 2 and "decimate 1" in e.value.msg   # the reference's own assert location (Filter.hs:586)


def test_fm_chain_connected_pipes_vs_oracle(sdr, port):
    """cfg4: u8 IQ >-> decimate-by-8 >-> fmDemod >-> resample 3/10 >-> 64-tap filter >-> (*0.2)   (fm.hs:34-41)"""
    ctx = sdr.default_context()
    n_bufs, buf = 24, 16384   # 8192 IQ pairs per buffer
    raw = synth.rand_bytes(n_bufs * buf)
    t_dec = synth.windowed_sinc_taps(128, 1 / 16)
    t_res = synth.windowed_sinc_taps(90, 1 / 20, gain=3.0)
    half = synth.windowed_sinc_taps(64, 1 / 4)[:32]
    dec = sdr.cudaDecimatorC(8, t_dec, sizeMultiple=4)
    res = sdr.cudaResamplerR(3, 10, t_res, sizeMultiple=8)
    fil = sdr.cudaFilterSymR(half)
    p0 = sdr.pipeConvertU8(ctx)
    p1 = sdr.pipeFirDecimator(dec, 8192)
    p2 = sdr.pipeFmDemod(ctx)
    p3 = sdr.pipeFirResampler(res, 1024)
    p4 = sdr.pipeFirFilter(fil, 1024)
    p5 = sdr.pipeScale(0.2, ctx)
    p0.connect(p1).connect(p2).connect(p3).connect(p4).connect(p5)
    got = []
    for i in range(n_bufs):
        p0.push(raw[i * buf:(i + 1) * buf])
        while p5.ready():
            got.append(p5.pop(1024))
    # oracle chain
    bufs = [raw[i * buf:(i + 1) * buf] for i in range(n_bufs)]
    s0 = (port.convert_u8(b).view(np.complex64) for b in bufs)
    s1 = op.fir_decimator(op.mk_decimator_c(V_AVX, 8, t_dec), 8192, s0)
    s2 = op.fm_demod(s1)
    s3 = op.fir_resampler(op.mk_resampler(V_AVX, 3, 10, t_res), 1024, s2)
    s4 = op.fir_filter(op.mk_filter_sym_r(V_AVXSYM, half), 1024, s3)
    want = [port.scale(np.float32(0.2), v) for v in s4]
    assert len(got) == len(want) and len(got) >= 1
    g, w = np.concatenate(got), np.concatenate(want)
    # fmDemod sits in the middle: atan2 of near-zero products amplifies upstream rounding, so this end-to-end check
    # is on the output scale only
    assert np.abs(g - w).max() <= 1e-4 * max(1.0, float(np.abs(w).max()))


@pytest.mark.parametrize("block_out,sizes", [(8192, [16384 * 8] * 6), (1000, [16384, 2 * 3001, 2 * 300, 2 * 9000, 2 * 40000, 2 * 70000]),
                                             (4096, [1 << 22, 1 << 22])])
@pytest.mark.parametrize("symmetric", [True, False])
def test_fused_fm_frontend_equals_unfused_chain(sdr, port, block_out, sizes, symmetric):
    """sdr_pipe_fm_frontend == convert >-> firDecimator >-> fmDemod connected stage by stage, BIT FOR BIT (same FIR
    summation order, same discriminator), and both match the oracle chain"""
    ctx = sdr.default_context()
    raw = synth.rand_bytes(sum(sizes))
    t_dec = synth.windowed_sinc_taps(128, 1 / 16)
    if not symmetric:   # arbitrary taps take the 8-warp kernel that keeps all 128 in registers
        t_dec = (t_dec * np.linspace(0.5, 1.5, 128)).astype(np.float32)
    dec = sdr.cudaDecimatorC(8, t_dec, sizeMultiple=4)
    fused = sdr.pipeFmFrontEnd(dec, block_out)
    p0 = sdr.pipeConvertU8(ctx)
    p1 = sdr.pipeFirDecimator(dec, block_out)
    p2 = sdr.pipeFmDemod(ctx)
    p0.connect(p1).connect(p2)
    a, b, i = [], [], 0
    used_fused = False
    for n in sizes:
        fused.push(raw[i:i + n])
        p0.push(raw[i:i + n])
        i += n
        while fused.ready():
            a.append(fused.pop())
        used_fused |= sdr._lib.lib.sdr_pipe_last_kernel(fused.h).decode().startswith("fm_front_ring")
        while p2.ready():
            b.append(p2.pop())
    assert len(a) == len(b) and len(a) >= 1 and all(len(v) == block_out for v in a)
    assert used_fused
    assert ("sym" in sdr._lib.lib.sdr_pipe_last_kernel(fused.h).decode()) == symmetric
    A, B = np.concatenate(a), np.concatenate(b)
    assert np.array_equal(A, B), f"{int((A != B).sum())} of {len(A)} differ, first at {int(np.argmax(A != B))}"
    # oracle chain on a prefix (CPU cost)
    npre = min(len(raw), 16384 * 12)
    x = port.convert_u8(raw[:npre]).view(np.complex64)
    y = port.decimate(V_AVX, (len(x) - 128) // 8 + 1, 8, np.repeat(t_dec, 2), x, True)
    want = port.fm_demod(y, 0j)
    m = min(len(want), len(A))
    assert np.abs(A[:m] - want[:m]).max() <= 2e-4   # atan2 of small products amplifies the 1e-7 FIR rounding difference


# ---------------------------------------------------------------------------------------------------------------
# full-size, size-independent properties on device-resident streams
# ---------------------------------------------------------------------------------------------------------------
def test_synth_noise_matches_cpu_replica(sdr, ctx):
    n = 1 << 16
    buf = ctx.alloc(4 * n)
    ctx.synth_noise(buf, n, first_float=12345)
    assert np.array_equal(buf.to_host(np.float32, n), synth.noise(n, first=12345))
    ctx.synth_bytes(buf, n, first_byte=77)
    assert np.array_equal(buf.to_host(np.uint8, n), synth.rand_bytes(n, first=77))
    ctx.synth_noise(buf, n, first_float=0)
    assert ctx.checksum32(buf, n, first_word=5) == synth.checksum32(synth.noise(n), first=5)
    buf.free()


@pytest.mark.parametrize("log2n", [17, 24, 27])
def test_stream_decimator_tuned_vs_generic_and_oracle(sdr, L, ctx, ref, log2n):
    """Full-size parity through size-independent properties: (1) the tuned ring kernel and the generic kernel sum
    in the same order, so the checksum of their outputs over the whole stream must be IDENTICAL; (2) sampled windows of
    the big stream are compared with the reference AVX C; (3) linearity in the taps."""
    n = 1 << log2n
    taps = synth.windowed_sinc_taps(128, 1 / 16)
    d = sdr.cudaDecimatorC(8, taps, sizeMultiple=4)
    num = (n - 128) // 8 + 1
    x = ctx.alloc(8 * n + 64)
    y = ctx.alloc(8 * num + 64)
    y2 = ctx.alloc(8 * num + 64)
    ctx.synth_noise(x, 2 * n)
    L.check(L.lib.sdr_decimate_stream(d.handle, x.ptr, n, y.ptr, num))
    assert d.last_kernel().startswith("dec_c_ring")
    # (1) generic path: same stream shifted by one sample (8 B) so the input is not 16-byte aligned
    ctx.synth_noise(x, 2 * n, first_float=0, offset_bytes=8)
    L.check(L.lib.sdr_decimate_stream(d.handle, x.at(8), n, y2.ptr, num))
    assert d.last_kernel() == "fir_direct"
    assert ctx.checksum32(y, 2 * num) == ctx.checksum32(y2, 2 * num)
    # (2) windows of the stream against the reference C (regenerated on the CPU from the counter RNG)
    for m0 in sorted({0, 255, 256, 1009, num // 3, num - 700}):
        m0 = max(0, min(m0, num - 600))
        cnt = 600
        xs = synth.noise_complex(cnt * 8 + 128, first=m0 * 8)
        want = ref.decimate("decimateAVXRC", cnt, 8, np.repeat(taps, 2), xs)
        got = y.to_host(np.complex64, cnt, offset_bytes=8 * m0)
        close(got, want)
    # (3) linearity: taps -> 2 * taps doubles every output exactly (power-of-two scaling is exact in binary32)
    d2 = sdr.cudaDecimatorC(8, taps * np.float32(2), sizeMultiple=4)
    L.check(L.lib.sdr_decimate_stream(d2.handle, x.at(8), n, y2.ptr, num))
    a = y.to_host(np.complex64, 4096, offset_bytes=8 * (num // 2))
    b = y2.to_host(np.complex64, 4096, offset_bytes=8 * (num // 2))
    assert np.array_equal(a * np.complex64(2), b)
    for b_ in (x, y, y2):
        b_.free()


@pytest.mark.parametrize("T", [64, 32, 128])
@pytest.mark.parametrize("log2n", [18, 26])
def test_stream_real_filter_tuned_vs_generic_and_oracle(sdr, L, ctx, ref, T, log2n):
    """cfg1 at full size: ring kernel == generic kernel bit for bit (same summation order), windows vs reference AVX"""
    n = 1 << log2n
    half = synth.windowed_sinc_taps(T, 1 / 4)[:T // 2]
    f = sdr.cudaFilterSymR(half)
    num = n - T + 1
    x = ctx.alloc(4 * n + 64)
    y = ctx.alloc(4 * num + 64)
    y2 = ctx.alloc(4 * num + 64)
    ctx.synth_noise(x, n)
    L.check(L.lib.sdr_filter_stream(f.handle, x.ptr, n, y.ptr, num))
    assert f.last_kernel().startswith("fir_r_ring"), f.last_kernel()
    ctx.synth_noise(x, n, first_float=0, offset_bytes=4)   # same stream, 4-byte shifted: not 16-byte aligned -> generic
    L.check(L.lib.sdr_filter_stream(f.handle, x.at(4), n, y2.ptr, num))
    assert not f.last_kernel().startswith("fir_r_ring")
    assert ctx.checksum32(y, num) == ctx.checksum32(y2, num)
    for m0 in sorted({0, 3839, 3840, num // 2, num - 4000}):
        cnt = 2048
        xs = synth.noise(cnt + T, first=m0)
        want = ref.filter("filterAVXSymmetricRR", cnt, half, xs)
        close(y.to_host(np.float32, cnt, offset_bytes=4 * m0), want)
    for b_ in (x, y, y2):
        b_.free()


@pytest.mark.parametrize("T", [64, 32, 128])
@pytest.mark.parametrize("log2n", [18, 26])
def test_stream_real_filter_fast_fir_vs_direct_and_oracle(sdr, L, ctx, ref, T, log2n):
    """cfg1 with the opt-in 2-parallel fast-FIR arithmetic (sdr_ctx_set_fast_fir): same ring, sums associated differently,
    so parity is the path's tolerance against the reference AVX filter AND against the direct-form ring kernel"""
    n = 1 << log2n
    half = synth.windowed_sinc_taps(T, 1 / 4)[:T // 2]
    f = sdr.cudaFilterSymR(half)
    num = n - T + 1
    x, y, y2 = ctx.alloc(4 * n + 64), ctx.alloc(4 * num + 64), ctx.alloc(4 * num + 64)
    ctx.synth_noise(x, n)
    try:
        L.check(L.lib.sdr_filter_stream(f.handle, x.ptr, n, y2.ptr, num))
        assert f.last_kernel().startswith("fir_r_ring"), f.last_kernel()
        ctx.set_fast_fir(True)
        L.check(L.lib.sdr_filter_stream(f.handle, x.ptr, n, y.ptr, num))
        assert f.last_kernel().startswith("fir_r_ffa_ring"), f.last_kernel()
        for m0 in sorted({0, 3839, 3840, num // 2, num - 4000}):
            cnt = 2048
            xs = synth.noise(cnt + T, first=m0)
            want = ref.filter("filterAVXSymmetricRR", cnt, half, xs)
            got = y.to_host(np.float32, cnt, offset_bytes=4 * m0)
            close(got, want)
            close(got, y2.to_host(np.float32, cnt, offset_bytes=4 * m0))
        # the ragged end is finished by the generic kernel either way: the last outputs agree bit for bit
        assert np.array_equal(y.to_host(np.float32, 8, offset_bytes=4 * (num - 8)), y2.to_host(np.float32, 8, offset_bytes=4 * (num - 8)))
    finally:
        ctx.set_fast_fir(False)
        for b_ in (x, y, y2):
            b_.free()


@pytest.mark.parametrize("T", [90, 31])
@pytest.mark.parametrize("log2n", [18, 26])
def test_stream_real_resampler_tuned_vs_generic_and_oracle(sdr, L, ctx, ref, T, log2n):
    """cfg3 at full size: ring kernel == generic kernel bit for bit; windows vs the reference's resampleAVXRR"""
    n = 1 << log2n
    taps = synth.windowed_sinc_taps(T, 1 / 20, gain=3.0)
    r = sdr.cudaResamplerR(3, 10, taps, sizeMultiple=8)
    num = (n * 3 - r.numCoeffsR) // 10 + 1
    x = ctx.alloc(4 * n + 64)
    y = ctx.alloc(4 * num + 64)
    y2 = ctx.alloc(4 * num + 64)
    ctx.synth_noise(x, n)
    L.check(L.lib.sdr_resample_stream(r.handle, x.ptr, n, y.ptr, num))
    assert r.last_kernel().startswith("res_r_ring"), r.last_kernel()
    ctx.synth_noise(x, n, first_float=0, offset_bytes=4)
    # misaligned start: the tuned kernel only engages after a generic prefix reaches an aligned cycle boundary
    L.check(L.lib.sdr_resample_stream(r.handle, x.at(4), n, y2.ptr, num))
    assert ctx.checksum32(y, num) == ctx.checksum32(y2, num)
    num_coeffs, increments, groups = op.prepare_coeffs(8, 3, 10, taps)
    for k0 in sorted({0, 3 * 1000, 3 * (num // 6), 3 * ((num - 3000) // 3)}):
        cnt = 1500
        i0 = (k0 * 10 + 2) // 3
        xs = synth.noise(cnt * 10 // 3 + 200, first=i0)
        want, _ = ref.resample("resampleAVXRR", cnt, num_coeffs, 0, increments, groups, xs)
        close(y.to_host(np.float32, cnt, offset_bytes=4 * k0), want)
    for b_ in (x, y, y2):
        b_.free()


def test_sharded_plan_single_rank_equals_stream(sdr, L, ctx):
    """world = 1 plan through sdr_decimate_sharded == sdr_decimate_stream (no communicator needed)"""
    n = 1 << 20
    taps = synth.windowed_sinc_taps(128, 1 / 16)
    d = sdr.cudaDecimatorC(8, taps, sizeMultiple=4)
    plan = sdr.multigpu.shard_plan(n, 128, 8, 1, 0)
    x = ctx.alloc(8 * n)
    y = ctx.alloc(8 * plan.out_count)
    y2 = ctx.alloc(8 * plan.out_count)
    ctx.synth_noise(x, 2 * n)
    sdr.multigpu.decimate_sharded(d, None, plan, x.ptr, y.ptr)
    L.check(L.lib.sdr_decimate_stream(d.handle, x.ptr, n, y2.ptr, plan.out_count))
    assert ctx.checksum32(y, 2 * plan.out_count) == ctx.checksum32(y2, 2 * plan.out_count)


@pytest.mark.parametrize("n_last,n_next,count", [
    (2048 * 5, 120, 2048 * 5 // 8),              # a shard boundary: every window that STARTS in `last`, 120-sample halo
    ((1 << 16) + 2, 1000, None),                 # boundary inside a lane segment, ragged last sub-tile
    ((1 << 16) + 64 * 3, 128, None),             # boundary on a lane-segment edge
    (4096, 4096, 300),                           # only the head of `next` is needed
])
def test_ring_kernel_covers_two_segments_and_ragged_end(sdr, L, ctx, ref, n_last, n_next, count):
    """the headline ring kernel in COVERING mode (one launch, no generic tail): windows that straddle lastBuf ++ nextBuf
    (what a sharded pass reads from its neighbour) and the ragged last sub-tile, against the reference AVX C"""
    taps = synth.windowed_sinc_taps(128, 1 / 16)
    d = sdr.cudaDecimatorC(8, taps, sizeMultiple=4)
    if count is None:
        count = (n_last + n_next - 128) // 8 + 1
    xs = synth.noise_complex(n_last + n_next, first=12345)
    a = ctx.to_device(xs[:n_last])
    b = ctx.to_device(xs[n_last:])
    y = ctx.alloc(8 * count + 64)
    L.check(L.lib.sdr_memset_dev(ctx.h, y.ptr, 0xff, 8 * count + 64))
    before = ctx.launches
    L.check(L.lib.sdr_decimate_cross(d.handle, count, a.ptr, n_last, b.ptr, n_next, y.ptr, L.SDR_DEVICE))
    assert d.last_kernel().startswith("dec_c_ring") and ctx.launches - before == 1, (d.last_kernel(), ctx.launches - before)
    got = y.to_host(np.complex64, count + 8)
    want = ref.decimate("decimateAVXRC", count, 8, np.repeat(taps, 2), xs)
    close(got[:count], want)
    assert np.all(got[count:].view(np.uint32) == 0xffffffff), "stores past the last output"
    for buf in (a, b, y):
        buf.free()


@pytest.mark.parametrize("cplx", [True, False])
def test_decimator_pipe_factor_larger_than_tap_count(sdr, cplx):
    """decimation factor > numCoeffs: the next window starts beyond the resident samples.  The reference runs into its own
    `decimate 3` assert there (Filter.hs:603); the pipe carries the samples still to skip into the following vectors and
    keeps computing the flat-stream decimation y[m] = sum_k c[k] x[m D + k]."""
    T, D, block_out = 16, 40, 50
    sizes = [4096, 1000, 40, 17, 5000, 16, 3000]
    taps = taps_for(T, 3)
    x = rnd(sum(sizes), cplx, 8)
    d = (sdr.cudaDecimatorC if cplx else sdr.cudaDecimatorR)(D, taps)
    got = list(sdr.firDecimator(d, block_out, _chunks(x, sizes)))
    total = (len(x) - T) // D + 1
    assert len(got) == total // block_out and all(len(g) == block_out for g in got)
    y = np.concatenate(got)
    close(y, op.flat_decimate(x, taps, D, len(y)).astype(y.dtype))
