"""Zero-copy device pushes (SDR_DEVICE_HELD) and the fused low-rate stage.

* A FIR stage fed with device vectors it may keep referring to (the Pipes contract: a yielded vector is immutable,
  Filter.hs:519-521) reads them in place; the stream it yields must be bit-identical to the same vectors pushed from host
  memory, for every stage kind, ragged vector sizes, adjacent and non-adjacent vectors.
* `firResampler >-> firFilter >-> P.map (* k)` fused (sdr_pipe_fm_lowrate, fm.hs:38-40) == the three stages connected one
  after the other, bit for bit; the whole FM chain as two fused stages == the six un-fused stages."""
import ctypes as C
import os

import numpy as np
import pytest

import synth

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
FM = np.load(os.path.join(HERE, "golden", "fm_example_coeffs.npz"))


@pytest.fixture(scope="module")
def sdr():
    import sdr_b200
    assert sdr_b200.has_cuda()
    return sdr_b200


@pytest.fixture(scope="module")
def ctx(sdr):
    return sdr.default_context()


def _noise(n, cplx, seed):
    rng = np.random.default_rng(seed)
    if cplx:
        return (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    return rng.standard_normal(n).astype(np.float32)


def _drain(pipe, out):
    while pipe.ready():
        out.append(pipe.pop())


def _makers(sdr):
    taps128 = synth.windowed_sinc_taps(128, 1 / 16)
    half = synth.windowed_sinc_taps(64, 1 / 4)[:32]
    t90 = synth.windowed_sinc_taps(90, 1 / 20, gain=3.0)
    return {
        "decimatorC_128_8": (lambda: sdr.pipeFirDecimator(sdr.cudaDecimatorC(8, taps128, sizeMultiple=4), 1024), lambda n: _noise(n, True, 1), "ring"),
        "decimatorC_51_8": (lambda: sdr.pipeFirDecimator(sdr.cudaDecimatorC(8, FM["coeffsRFDecim"], sizeMultiple=4), 1024), lambda n: _noise(n, True, 2), "ring"),
        "decimatorR_64_4": (lambda: sdr.pipeFirDecimator(sdr.cudaDecimatorR(4, synth.windowed_sinc_taps(64, 1 / 8), sizeMultiple=8), 1024), lambda n: _noise(n, False, 3), "ring"),
        "filterR_64": (lambda: sdr.pipeFirFilter(sdr.cudaFilterSymR(half), 4096), lambda n: _noise(n, False, 4), "ring"),
        "filterC_32": (lambda: sdr.pipeFirFilter(sdr.cudaFilterC(synth.windowed_sinc_taps(32, 1 / 4), sizeMultiple=8), 4096), lambda n: _noise(n, True, 5), "ring"),
        "resamplerR_90": (lambda: sdr.pipeFirResampler(sdr.cudaResamplerR(3, 10, t90, sizeMultiple=8), 2048), lambda n: _noise(n, False, 6), "ring"),
        "resamplerC_90": (lambda: sdr.pipeFirResampler(sdr.cudaResamplerC(3, 10, t90, sizeMultiple=8), 2048), lambda n: _noise(n, True, 7), "ring"),
        "fmFrontEnd": (lambda: sdr.pipeFmFrontEnd(sdr.cudaDecimatorC(8, taps128, sizeMultiple=4), 1024),
                       lambda n: np.random.default_rng(8).integers(0, 256, 2 * n, dtype=np.uint8), "fm_front_ring"),
        "u8Decimator": (lambda: sdr.pipeU8Decimator(sdr.cudaDecimatorC(8, taps128, sizeMultiple=4), 1024),
                        lambda n: np.random.default_rng(9).integers(0, 256, 2 * n, dtype=np.uint8), "dec_u8_ring"),
    }


SIZES = [262144, 8192, 8192, 3001 * 2, 131072 + 6, 8192, 70000, 262144 + 2]


@pytest.mark.parametrize("kind", ["decimatorC_128_8", "decimatorC_51_8", "decimatorR_64_4", "filterR_64", "filterC_32", "resamplerR_90",
                                  "resamplerC_90", "fmFrontEnd", "u8Decimator"])
@pytest.mark.parametrize("layout", ["adjacent", "scattered"])
def test_held_device_vectors_equal_host_pushes(sdr, ctx, kind, layout):
    make, data, ring_name = _makers(sdr)[kind]
    byte_fed = kind in ("fmFrontEnd", "u8Decimator")
    k = 2 if byte_fed else 1
    x = data(sum(SIZES))
    # host pushes (staged copies)
    p = make()
    want, o = [], 0
    for n in SIZES:
        p.push(x[k * o:k * (o + n)])
        _drain(p, want)
        o += n
    p.close()
    # the same vectors resident on the device, read in place
    item = x.dtype.itemsize
    gap = 4096 if layout == "scattered" else 0           # bytes between consecutive vectors
    dbuf = ctx.alloc(x.nbytes + gap * len(SIZES) + 256)
    offs, o, pos = [], 0, 0
    for n in SIZES:
        dbuf.upload(x[k * o:k * (o + n)], offset_bytes=pos)
        offs.append(pos)
        pos += k * n * item + gap
        o += n
    p = make()
    got, used_ring = [], False
    for n, off in zip(SIZES, offs):
        p.push_device(dbuf.at(off), k * n, held=True)
        used_ring |= ring_name in sdr._lib.lib.sdr_pipe_last_kernel(p.h).decode() or ring_name in (p.owner.last_kernel() if hasattr(p.owner, "last_kernel") else "")
        _drain(p, got)
    p.sync()
    _drain(p, got)
    p.close()
    dbuf.free()
    assert used_ring, "the in-place run never reached a tuned kernel"
    assert len(got) == len(want) and len(want) > 0
    for a, b in zip(got, want):
        assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), kind


@pytest.mark.parametrize("taps_r,L,M,expect", [("t90", 3, 10, "fm_lowrate<3,10,90,64>"), ("fm31", 3, 10, "fm_lowrate<3,10,31,64>"), ("t40", 2, 7, "unfused")])
@pytest.mark.parametrize("sizes", [[8192] * 12, [50000, 8192, 130001, 3000, 8192, 99999]])
def test_fused_lowrate_stage_equals_the_three_stages(sdr, taps_r, L, M, expect, sizes):
    tr = {"t90": synth.windowed_sinc_taps(90, 1 / 20, gain=3.0), "fm31": FM["coeffsAudioResampler"],
          "t40": synth.windowed_sinc_taps(40, 1 / 14, gain=2.0)}[taps_r]
    half = FM["coeffsAudioFilter"]
    x = _noise(sum(sizes), False, 11)
    vecs, o = [], 0
    for n in sizes:
        vecs.append(x[o:o + n]); o += n
    br, bf = 2048, 1000
    r = sdr.cudaResamplerR(L, M, tr, sizeMultiple=8)
    f = sdr.cudaFilterSymR(half)
    fused = sdr.pipeFmLowRate(r, br, f, bf, 0.2)
    got = []
    for v in vecs:
        fused.push(v)
        _drain(fused, got)
    assert sdr._lib.lib.sdr_pipe_last_kernel(fused.h).decode() == expect
    p1, p2, p3 = sdr.pipeFirResampler(r, br), sdr.pipeFirFilter(f, bf), sdr.pipeScale(0.2)
    p1.connect(p2).connect(p3)
    want = []
    for v in vecs:
        p1.push(v)
        _drain(p3, want)
    assert len(got) == len(want) and len(want) > 3, (len(got), len(want))
    for a, b in zip(got, want):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    for p in (fused, p1, p2, p3):
        p.close()


def test_fm_chain_two_fused_stages_equal_six_stages(sdr, ctx):
    """u8 IQ -> [convert + decimate + fmDemod] -> [resample + filter + scale] (two launches per push) == the six stages"""
    taps = synth.windowed_sinc_taps(128, 1 / 16)
    t90 = synth.windowed_sinc_taps(90, 1 / 20, gain=3.0)
    d = sdr.cudaDecimatorC(8, taps, sizeMultiple=4)
    r = sdr.cudaResamplerR(3, 10, t90, sizeMultiple=8)
    f = sdr.cudaFilterSymR(FM["coeffsAudioFilter"])
    rng = np.random.default_rng(3)
    raw = rng.integers(0, 256, 2 * (1 << 22), dtype=np.uint8)
    sizes = [1 << 21, 16384, 16384, 1 << 22, 2 * 3000, (1 << 21) - 2 * 3000 - 32768]
    vecs, o = [], 0
    for s in sizes:
        vecs.append(raw[o:o + s]); o += s

    def run(head, tail):
        out = []
        for v in vecs:
            head.push(v)
            _drain(tail, out)
        return out

    fe, lo = sdr.pipeFmFrontEnd(d, 8192), sdr.pipeFmLowRate(r, 8192, f, 8192, 0.2)
    fe.connect(lo)
    a = run(fe, lo)
    q = [sdr.pipeConvertU8(), sdr.pipeFirDecimator(d, 8192), sdr.pipeFmDemod(), sdr.pipeFirResampler(r, 8192), sdr.pipeFirFilter(f, 8192),
         sdr.pipeScale(0.2)]
    for s0, s1 in zip(q, q[1:]):
        s0.connect(s1)
    b = run(q[0], q[-1])
    assert len(a) == len(b) and len(a) >= 15, (len(a), len(b))
    for u, v in zip(a, b):
        assert np.array_equal(u.view(np.uint32), v.view(np.uint32))
    for p in [fe, lo] + q:
        p.close()


@pytest.mark.parametrize("threshold,order", [(1 << 14, "down_first"), (1 << 18, "up_first"), (1 << 23, "down_first")])
def test_connected_stage_keeps_handed_over_vectors_in_the_upstream_fifo(sdr, ctx, threshold, order):
    """a launch threshold on the downstream stage (sdr_pipe_set_batch) makes hand-overs accumulate: they stay in the upstream
    stage's FIFO (no copy), which must then neither rewind nor move under them -- until it runs out of room (with the 2^23
    threshold nothing launches before the end of the input: the upstream FIFO has to grow several times under the run).
    Same stream as the un-thresholded stages given everything at once; either stage may be destroyed first."""
    from sdr_b200 import _lib as L
    taps = synth.windowed_sinc_taps(128, 1 / 16)
    t90 = synth.windowed_sinc_taps(90, 1 / 20, gain=3.0)
    d = sdr.cudaDecimatorC(8, taps, sizeMultiple=4)
    r = sdr.cudaResamplerR(3, 10, t90, sizeMultiple=8)
    f = sdr.cudaFilterSymR(FM["coeffsAudioFilter"])
    vec, nvec = 1 << 20, 15
    raw = np.random.default_rng(11).integers(0, 256, vec * nvec, dtype=np.uint8)
    ref = []
    stages = [sdr.pipeFmFrontEnd(d, 8192), sdr.pipeFmLowRate(r, 8192, f, 8192, 0.2)]
    stages[0].connect(stages[1])
    stages[0].push(raw)
    stages[1].sync()
    _drain(stages[1], ref)
    for p in stages:
        p.close()
    ref = np.concatenate(ref)
    fe, lo = sdr.pipeFmFrontEnd(d, 8192), sdr.pipeFmLowRate(r, 8192, f, 8192, 0.2)
    fe.connect(lo)
    L.check(L.lib.sdr_pipe_set_batch(lo.h, threshold))
    dbuf = ctx.to_device(raw)
    out = ctx.alloc(4 * len(ref) + 65536)
    n_out = C.c_longlong()
    L.check(L.lib.sdr_pipe_run(fe.h, lo.h, dbuf.ptr, vec, nvec, L.SDR_DEVICE_HELD, out.ptr, len(ref) + 8192, L.SDR_DEVICE, C.byref(n_out)))
    assert n_out.value == len(ref) and len(ref) > 100000
    got = out.to_host(np.float32, len(ref))
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    for p in ((lo, fe) if order == "down_first" else (fe, lo)):
        p.close()
    dbuf.free(); out.free()


@pytest.mark.parametrize("slack", [8192, 0, -1])
def test_pipe_run_produces_straight_into_a_device_output_buffer(sdr, ctx, slack):
    """sdr_pipe_run with a device output buffer: the sink stage works IN the caller's buffer (no copy of the yielded
    vectors).  Two runs through one stage -- the partial output block the first run leaves goes in front in the second --
    with ample capacity, with exactly the capacity the yielded vectors need (the partial block does not fit: the stage takes
    its FIFO back mid-run and copies from there on) and one element less (an error, not an overrun).  Same vectors as
    host pushes / pops."""
    from sdr_b200 import _lib as L
    taps = synth.windowed_sinc_taps(128, 1 / 16)
    vec, nvec = 8192 + 6, 37
    x = synth.noise_complex(vec * nvec * 2, first=21)
    d = sdr.cudaDecimatorC(8, taps, sizeMultiple=4)
    ref_pipe = sdr.pipeFirDecimator(d, 1000)
    ref = [[], []]
    for run in range(2):
        for i in range(nvec):
            ref_pipe.push(x[(run * nvec + i) * vec:(run * nvec + i + 1) * vec])
            _drain(ref_pipe, ref[run])
    ref_pipe.close()
    p = sdr.pipeFirDecimator(d, 1000)
    dbuf = ctx.to_device(x)
    n_out = C.c_longlong()
    for run in range(2):
        want = np.concatenate(ref[run])
        cap = len(want) + slack
        out = ctx.alloc(8 * (len(want) + 8192) + 64)
        L.check(L.lib.sdr_memset_dev(ctx.h, out.ptr, 0xff, 8 * (len(want) + 8192)))
        rc = L.lib.sdr_pipe_run(p.h, p.h, dbuf.at(8 * run * nvec * vec), vec, nvec, L.SDR_DEVICE_HELD, out.ptr, cap, L.SDR_DEVICE, C.byref(n_out))
        if slack < 0:
            assert rc != 0            # capacity too small for the yielded vectors
            out.free()
            break
        L.check(rc)
        assert n_out.value == len(want)
        got = out.to_host(np.complex64, len(want))
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
        if slack == 0:                # nothing may have been written behind the capacity
            tail = out.to_host(np.uint32, 16, offset_bytes=8 * cap)
            assert np.all(tail == 0xffffffff)
        out.free()
    p.close()
    dbuf.free()
