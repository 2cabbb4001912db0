#!/usr/bin/env python
"""One launch of each tuned kernel on a stream larger than L2 (for `ncu --set full`): see tools/gpu_r2_s9.sh."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import sdr_b200  # noqa: E402
from sdr_b200 import _lib as L  # noqa: E402


def main():
    log2 = int(sys.argv[1]) if len(sys.argv) > 1 else 27
    n = 1 << log2
    ctx = sdr_b200.default_context()
    x = ctx.alloc(8 * n + 256)
    y = ctx.alloc(8 * n + 256)
    ctx.synth_noise(x, 2 * n)
    fm = np.load(os.path.join(ROOT, "tests", "golden", "fm_example_coeffs.npz"))
    nr = 2 * n
    # cfg2 headline, FM example's 51-tap decimator, complex 64-tap filter, real decimator
    d = sdr_b200.cudaDecimatorC(8, sdr_b200.windowed_sinc_taps(128, 1 / 16), ctx=ctx, sizeMultiple=4)
    L.check(L.lib.sdr_decimate_stream(d.handle, x.ptr, n, y.ptr, (n - 128) // 8 + 1))
    d51 = sdr_b200.cudaDecimatorC(8, fm["coeffsRFDecim"], ctx=ctx, sizeMultiple=4)
    L.check(L.lib.sdr_decimate_stream(d51.handle, x.ptr, n, y.ptr, (n - 52) // 8 + 1))
    fc = sdr_b200.cudaFilterC(sdr_b200.windowed_sinc_taps(64, 1 / 4), ctx=ctx, sizeMultiple=8)
    L.check(L.lib.sdr_filter_stream(fc.handle, x.ptr, n, y.ptr, n - 63))
    dr = sdr_b200.cudaDecimatorR(8, sdr_b200.windowed_sinc_taps(128, 1 / 16), ctx=ctx, sizeMultiple=8)
    L.check(L.lib.sdr_decimate_stream(dr.handle, x.ptr, nr, y.ptr, (nr - 128) // 8 + 1))
    # 256 taps: the ring kernel with its taps as launch parameters
    d256 = sdr_b200.cudaDecimatorC(8, sdr_b200.windowed_sinc_taps(256, 1 / 16), ctx=ctx, sizeMultiple=4)
    L.check(L.lib.sdr_decimate_stream(d256.handle, x.ptr, n, y.ptr, (n - 256) // 8 + 1))
    # cfg1, cfg3, complex resampler
    half = sdr_b200.windowed_sinc_taps(64, 1 / 4)[:32]
    f = sdr_b200.cudaFilterSymR(half, ctx=ctx)
    L.check(L.lib.sdr_filter_stream(f.handle, x.ptr, nr, y.ptr, nr - 63))
    t90 = sdr_b200.windowed_sinc_taps(90, 1 / 20, gain=3.0)
    r = sdr_b200.cudaResamplerR(3, 10, t90, ctx=ctx, sizeMultiple=8)
    L.check(L.lib.sdr_resample_stream(r.handle, x.ptr, nr, y.ptr, (nr * 3 - r.numCoeffsR) // 10 + 1))
    rc = sdr_b200.cudaResamplerC(3, 10, t90, ctx=ctx, sizeMultiple=8)
    L.check(L.lib.sdr_resample_stream(rc.handle, x.ptr, n, y.ptr, (n * 3 - rc.numCoeffsR) // 10 + 1))
    # cfg4: the two fused stages, one 2^27-IQ-pair push read in place
    raw = ctx.alloc(2 * n + 256)
    ctx.synth_bytes(raw, 2 * n)
    fe, lo = sdr_b200.pipeFmFrontEnd(d, 8192), sdr_b200.pipeFmLowRate(r, 8192, f, 8192, 0.2)
    fe.connect(lo)
    n_out = C.c_longlong()
    L.check(L.lib.sdr_pipe_run(fe.h, lo.h, raw.ptr, 2 * n, 1, L.SDR_DEVICE_HELD, y.ptr, 2 * n, L.SDR_DEVICE, C.byref(n_out)))
    u8 = sdr_b200.pipeU8Decimator(d, 8192)
    L.check(L.lib.sdr_pipe_run(u8.h, u8.h, raw.ptr, 2 * n, 1, L.SDR_DEVICE_HELD, y.ptr, n, L.SDR_DEVICE, C.byref(n_out)))
    # dcBlocker
    d_fin = ctx.alloc(8)
    ctx.dc_blocker(x.ptr, y.ptr, nr, d_fin.ptr)
    ctx.sync()
    print("NCU_TARGETS_OK", n_out.value)


if __name__ == "__main__":
    main()
