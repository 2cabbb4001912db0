"""Tuned (shared-memory ring) kernels beyond the headline shape: every (data type, tap count, decimation) the dispatchers
serve, against the UNMODIFIED reference C (oracle/_ref) on the same seeded stream, and bit for bit against the generic
kernel (same summation order; the generic kernel is reached by shifting the stream off its 16-byte alignment).

Includes the reference's own FM receiver coefficient sets (examples/fm/Coeffs.hs: 51-tap RF decimator, 31-tap 3/10 audio
resampler, 32 half-taps = 64-tap audio filter; fixture tests/golden/fm_example_coeffs.npz made by make_fm_coeffs.py)."""
import os

import numpy as np
import pytest

import oracle.pipes as op
import synth

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def sdr():
    import sdr_b200
    assert sdr_b200.has_cuda()
    return sdr_b200


@pytest.fixture(scope="module")
def L():
    from sdr_b200 import _lib
    return _lib


@pytest.fixture(scope="module")
def ctx(sdr):
    return sdr.default_context()


def close(y, want, rtol=1e-5):
    scale = np.maximum(np.abs(want), np.sqrt(np.mean(np.abs(want) ** 2)))
    err = np.abs(np.asarray(y) - want)
    assert np.all(err <= rtol * scale), float((err / scale).max())


def taps_for(n, seed):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal(n) / np.sqrt(n)).astype(np.float32)


FM = np.load(os.path.join(HERE, "golden", "fm_example_coeffs.npz"))

SHAPES = [(cplx, T, D) for cplx in (True, False) for T in (32, 51, 64, 100, 128) for D in (1, 2, 4, 8, 16)]
# 129..256 taps, decimation 4 / 8 / 16: the ring kernels with their taps as launch parameters
SHAPES += [(cplx, T, D) for cplx in (True, False) for T in (256, 200) for D in (4, 8, 16)] + [(True, 132, 8)]


@pytest.mark.parametrize("cplx,T,D", SHAPES)
def test_ring_kernels_every_shape_vs_reference_and_generic(sdr, L, ctx, ref, cplx, T, D):
    n = (1 << 20) + 2 * 1234
    eb = 8 if cplx else 4
    sm = 8 if (D == 1 or not cplx) else 4         # the AVX constructors' padding (Filter.hs:169,284,324); filters: keep filterAVXRC in bounds
    taps = FM["coeffsRFDecim"] if T == 51 else taps_for(T, 7 * T + D)
    if D == 1:
        rec = (sdr.cudaFilterC if cplx else sdr.cudaFilterR)(taps, sizeMultiple=sm)
        Ts = rec.numCoeffsF
        run = lambda xin, nin, yout, num: L.check(L.lib.sdr_filter_stream(rec.handle, xin, nin, yout, num))
    else:
        rec = (sdr.cudaDecimatorC if cplx else sdr.cudaDecimatorR)(D, taps, sizeMultiple=sm)
        Ts = rec.numCoeffsD
        run = lambda xin, nin, yout, num: L.check(L.lib.sdr_decimate_stream(rec.handle, xin, nin, yout, num))
    assert Ts == -(-T // sm) * sm
    num = (n - Ts) // D + 1
    nf = n * (2 if cplx else 1)
    x = ctx.alloc(eb * n + 64)
    y = ctx.alloc(eb * num + 64)
    y2 = ctx.alloc(eb * num + 64)
    ctx.synth_noise(x, nf)
    run(x.ptr, n, y.ptr, num)
    assert "ring" in rec.last_kernel(), rec.last_kernel()
    # generic kernel: the same stream shifted by one element, so the input is not 16-byte aligned
    ctx.synth_noise(x, nf, first_float=0, offset_bytes=eb)
    run(x.at(eb), n, y2.ptr, num)
    assert "ring" not in rec.last_kernel(), rec.last_kernel()
    w = eb // 4
    assert ctx.checksum32(y, w * num) == ctx.checksum32(y2, w * num)
    # windows against the reference AVX C (taps zero-padded exactly as the constructor stores them)
    padded = np.zeros(Ts, np.float32)
    padded[:T] = taps
    for m0 in sorted({0, 255, num // 2, num - 700}):
        cnt = 600
        first = m0 * D
        xs = synth.noise_complex(cnt * D + Ts, first=first) if cplx else synth.noise(cnt * D + Ts, first=first)
        if D == 1:
            want = ref.filter("filterAVXRC", cnt, np.repeat(padded, 2), xs) if cplx else ref.filter("filterAVXRR", cnt, padded, xs)
        else:
            want = (ref.decimate("decimateAVXRC", cnt, D, np.repeat(padded, 2), xs) if cplx
                    else ref.decimate("decimateAVXRR", cnt, D, padded, xs))
        got = y.to_host(np.complex64 if cplx else np.float32, cnt, offset_bytes=eb * m0)
        close(got, want)
    for b in (x, y, y2):
        b.free()


@pytest.mark.parametrize("T", [90, 31])
def test_complex_resampler_ring_vs_reference_and_generic(sdr, L, ctx, ref, T):
    """fastResamplerC 3/10 (resampleAVXRC, resample.c:125-142) on the ring kernel; 31 taps = the FM example's resampler set"""
    n = (1 << 20) + 600
    taps = FM["coeffsAudioResampler"] if T == 31 else synth.windowed_sinc_taps(T, 1 / 20, gain=3.0)
    r = sdr.cudaResamplerC(3, 10, taps, sizeMultiple=8)   # fastResamplerAVXC = mkResamplerC 8 (Filter.hs:491)
    num = (n * 3 - r.numCoeffsR) // 10 + 1
    x = ctx.alloc(8 * n + 64)
    y = ctx.alloc(8 * num + 64)
    y2 = ctx.alloc(8 * num + 64)
    ctx.synth_noise(x, 2 * n)
    L.check(L.lib.sdr_resample_stream(r.handle, x.ptr, n, y.ptr, num))
    assert r.last_kernel().startswith("res_c_ring"), r.last_kernel()
    ctx.synth_noise(x, 2 * n, first_float=0, offset_bytes=8)
    L.check(L.lib.sdr_resample_stream(r.handle, x.at(8), n, y2.ptr, num))   # every cycle boundary 8 bytes off: generic kernel
    assert not r.last_kernel().startswith("res_c_ring"), r.last_kernel()
    assert ctx.checksum32(y, 2 * num) == ctx.checksum32(y2, 2 * num)
    num_coeffs, increments, groups = op.prepare_coeffs(8, 3, 10, taps)
    for k0 in sorted({0, 3 * 1000, 3 * (num // 6), 3 * ((num - 3000) // 3)}):
        cnt = 1500
        i0 = (k0 * 10 + 2) // 3
        xs = synth.noise_complex(cnt * 10 // 3 + 200, first=i0)
        want, _ = ref.resample("resampleAVXRC", cnt, num_coeffs, 0, increments, groups, xs)
        close(y.to_host(np.complex64, cnt, offset_bytes=8 * k0), want)
    for b in (x, y, y2):
        b.free()


def test_fm_example_coefficient_sets_run_on_ring_kernels(sdr, L, ctx, ref):
    """the reference's own FM receiver sets (examples/fm/fm.hs:30-32): 51-tap RF decimator by 8 (stored 52), 31-tap 3/10
    audio resampler, 64-tap symmetric audio filter -- all on tuned kernels, windows against the reference C"""
    n = 1 << 21
    d = sdr.cudaDecimatorC(8, FM["coeffsRFDecim"], sizeMultiple=4)
    r = sdr.cudaResamplerR(3, 10, FM["coeffsAudioResampler"], sizeMultiple=8)
    f = sdr.cudaFilterSymR(FM["coeffsAudioFilter"])
    assert (d.numCoeffsD, r.numCoeffsR, f.numCoeffsF) == (52, 48, 64)
    x = ctx.alloc(8 * n + 64)
    y = ctx.alloc(8 * n + 64)
    ctx.synth_noise(x, 2 * n)
    L.check(L.lib.sdr_decimate_stream(d.handle, x.ptr, n, y.ptr, (n - 52) // 8 + 1))
    assert d.last_kernel() == "dec_c_ring<64,8,8>", d.last_kernel()
    pad = np.zeros(52, np.float32); pad[:51] = FM["coeffsRFDecim"]
    want = ref.decimate("decimateAVXRC", 600, 8, np.repeat(pad, 2), synth.noise_complex(600 * 8 + 52, first=8 * 5000))
    close(y.to_host(np.complex64, 600, offset_bytes=8 * 5000), want)
    nr = 2 * n
    L.check(L.lib.sdr_resample_stream(r.handle, x.ptr, nr, y.ptr, (nr * 3 - 48) // 10 + 1))
    assert r.last_kernel() == "res_r_ring<3,10,31,6,2>", r.last_kernel()
    L.check(L.lib.sdr_filter_stream(f.handle, x.ptr, nr, y.ptr, nr - 63))
    assert f.last_kernel() == "fir_r_ring<64,20,6>", f.last_kernel()
    want = ref.filter("filterAVXSymmetricRR", 2048, FM["coeffsAudioFilter"], synth.noise(2048 + 64, first=7777))
    close(y.to_host(np.float32, 2048, offset_bytes=4 * 7777), want)
    x.free(); y.free()


@pytest.mark.parametrize("T", [51, 64, 20])
def test_fused_front_end_smaller_tap_counts_equal_unfused_chain(sdr, T):
    """the fused u8 front ends at 64 / 32 tap capacity (the FM example's 51-tap decimator among them) == the stages one
    after the other, bit for bit"""
    taps = FM["coeffsRFDecim"] if T == 51 else taps_for(T, T)
    rng = np.random.default_rng(T)
    sizes = [16384, 2 * 3001, 2 * 40000, 2 * 300, 2 * 70000]
    raw = rng.integers(0, 256, sum(sizes), dtype=np.uint8)
    vecs, o = [], 0
    for s in sizes:
        vecs.append(raw[o:o + s]); o += s
    d = sdr.cudaDecimatorC(8, taps, sizeMultiple=4)

    def run(head, tail):
        out = []
        for v in vecs:
            head.push(v)
            while tail.ready():
                out.append(tail.pop())
        return np.concatenate(out) if out else np.zeros(0)

    fe = sdr.pipeFmFrontEnd(d, 500)
    a = run(fe, fe)
    assert sdr._lib.lib.sdr_pipe_last_kernel(fe.h).decode().startswith("fm_front_ring<%d" % (64 if T > 32 else 32))
    p0, p1, p2 = sdr.pipeConvertU8(), sdr.pipeFirDecimator(d, 500), sdr.pipeFmDemod()
    p0.connect(p1).connect(p2)
    b = run(p0, p2)
    n = min(len(a), len(b))
    assert n > 10000 and np.array_equal(a[:n].view(np.uint32), b[:n].view(np.uint32))
    u8 = sdr.pipeU8Decimator(d, 500)
    c = run(u8, u8)
    q0, q1 = sdr.pipeConvertU8(), sdr.pipeFirDecimator(d, 500)
    q0.connect(q1)
    e = run(q0, q1)
    assert len(c) == len(e) and np.array_equal(c.view(np.uint32), e.view(np.uint32))
    for p in (fe, p0, p1, p2, u8, q0, q1):
        p.close()
