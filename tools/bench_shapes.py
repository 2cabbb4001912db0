#!/usr/bin/env python
"""Device-resident throughput of every tuned shape (measurement tool): filters / decimators, real and complex, tap counts
32 / 51 / 64 / 128, decimation 1 / 2 / 4 / 8 / 16, the complex resampler and the FM example's own coefficient sets.  One
JSON line per shape with its HBM fraction and FP32-pipe fraction (measured peaks: MEASURED_PEAKS.json hbm_gbs, 33.5 TFMA/s)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import sdr_b200  # noqa: E402
from sdr_b200 import _lib as L  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
FP32 = 33.5e12


def timed(ctx, fn, steps=6, warm=2):
    for _ in range(warm):
        fn()
    ctx.sync()
    e0, e1 = sdr_b200.Event(ctx), sdr_b200.Event(ctx)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    return e0.elapsed_ms(e1) / steps


def main():
    log2 = int(sys.argv[1]) if len(sys.argv) > 1 else 27
    n = 1 << log2                 # complex samples; the real kernels see 2n floats
    ctx = sdr_b200.default_context()
    x = ctx.alloc(8 * n + 256)
    y = ctx.alloc(8 * n + 256)
    ctx.synth_noise(x, 2 * n)
    rng = np.random.default_rng(1)
    fm = np.load(os.path.join(ROOT, "tests", "golden", "fm_example_coeffs.npz"))
    for cplx in (True, False):
        ne = n if cplx else 2 * n
        eb = 8 if cplx else 4
        for T in (32, 51, 64, 128, 256):
            for D in (1, 2, 4, 8, 16):
                if T == 256 and D < 4:
                    continue          # 129..256 taps: decimators only (taps as launch parameters)
                taps = fm["coeffsRFDecim"] if T == 51 else (rng.standard_normal(T) / 8).astype(np.float32)
                sm = 8 if (D == 1 or not cplx) else 4
                if D == 1:
                    rec = (sdr_b200.cudaFilterC if cplx else sdr_b200.cudaFilterR)(taps, ctx=ctx, sizeMultiple=sm)
                    Ts = rec.numCoeffsF
                    fn = lambda: L.check(L.lib.sdr_filter_stream(rec.handle, x.ptr, ne, y.ptr, ne - Ts + 1))
                else:
                    rec = (sdr_b200.cudaDecimatorC if cplx else sdr_b200.cudaDecimatorR)(D, taps, ctx=ctx, sizeMultiple=sm)
                    Ts = rec.numCoeffsD
                    fn = lambda: L.check(L.lib.sdr_decimate_stream(rec.handle, x.ptr, ne, y.ptr, (ne - Ts) // D + 1))
                ms = timed(ctx, fn)
                rate = ne / (ms * 1e-3)
                bps = eb + eb / D
                Tk = 32 if Ts <= 32 else 64 if Ts <= 64 else 128 if Ts <= 128 else 256
                fma = (2 if cplx else 1) * Tk / D      # lane-FMAs per input element the kernel issues (its tap capacity)
                print(json.dumps({"data": "complex" if cplx else "real", "taps": T, "stored": Ts, "D": D, "kernel": rec.last_kernel(),
                                  "ms": round(ms, 4), "Gsamples_per_s": round(rate / 1e9, 1), "hbm_frac": round(rate * bps / 1e9 / PEAK, 3),
                                  "fp32_frac": round(rate * fma / FP32, 3)}), flush=True)
    for T in (90, 31):
        taps = fm["coeffsAudioResampler"] if T == 31 else sdr_b200.windowed_sinc_taps(T, 1 / 20, gain=3.0)
        for cplx in (False, True):
            ne = n if cplx else 2 * n
            eb = 8 if cplx else 4
            r = (sdr_b200.cudaResamplerC if cplx else sdr_b200.cudaResamplerR)(3, 10, taps, ctx=ctx, sizeMultiple=4 if cplx else 8)
            num = (ne * 3 - r.numCoeffsR) // 10 + 1
            ms = timed(ctx, lambda: L.check(L.lib.sdr_resample_stream(r.handle, x.ptr, ne, y.ptr, num)))
            rate = ne / (ms * 1e-3)
            print(json.dumps({"data": "complex" if cplx else "real", "resampler": "3/10", "taps": T, "kernel": r.last_kernel(), "ms": round(ms, 4),
                              "Gsamples_per_s": round(rate / 1e9, 1), "hbm_frac": round(rate * eb * 1.3 / 1e9 / PEAK, 3),
                              "fp32_frac": round(rate * (2 if cplx else 1) * ((T + 2) // 3) * 0.3 / FP32, 3)}), flush=True)


if __name__ == "__main__":
    main()
