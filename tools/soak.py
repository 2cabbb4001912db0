#!/usr/bin/env python
"""Soak: N passes of the ring decimator and of the ring resampler over the same stream with a checksum after every pass.
The hand-rolled mbarrier generation guard (csrc/ring_common.cuh) is the thing to break: a stale-phase read shows up as a
checksum that differs from the first pass."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sdr_b200  # noqa: E402
from sdr_b200 import _lib as L  # noqa: E402


def main():
    passes = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    log2 = int(sys.argv[2]) if len(sys.argv) > 2 else 26
    n = 1 << log2
    ctx = sdr_b200.default_context()
    x = ctx.alloc(8 * n + 256)
    y = ctx.alloc(8 * n + 256)
    ctx.synth_noise(x, 2 * n)
    d = sdr_b200.cudaDecimatorC(8, sdr_b200.windowed_sinc_taps(128, 1 / 16), ctx=ctx, sizeMultiple=4)
    r = sdr_b200.cudaResamplerR(3, 10, sdr_b200.windowed_sinc_taps(90, 1 / 20, gain=3.0), ctx=ctx, sizeMultiple=8)
    f = sdr_b200.cudaFilterSymR(sdr_b200.windowed_sinc_taps(64, 1 / 4)[:32], ctx=ctx)
    nr = 2 * n
    cases = {
        "dec_c_ring": (lambda: L.check(L.lib.sdr_decimate_stream(d.handle, x.ptr, n, y.ptr, (n - 128) // 8 + 1)), 2 * ((n - 128) // 8 + 1), d),
        "res_r_ring": (lambda: L.check(L.lib.sdr_resample_stream(r.handle, x.ptr, nr, y.ptr, (nr * 3 - r.numCoeffsR) // 10 + 1)),
                       (nr * 3 - r.numCoeffsR) // 10 + 1, r),
        "fir_r_ring": (lambda: L.check(L.lib.sdr_filter_stream(f.handle, x.ptr, nr, y.ptr, nr - 63)), nr - 63, f),
    }
    out = {}
    for name, (fn, nwords, rec) in cases.items():
        t0 = time.time()
        fn()
        first = ctx.checksum32(y, nwords)
        bad = 0
        for i in range(passes):
            L.check(L.lib.sdr_memset_dev(ctx.h, y.ptr, 0, 64))
            fn()
            if ctx.checksum32(y, nwords) != first:
                bad += 1
        out[name] = {"kernel": rec.last_kernel(), "passes": passes, "samples_per_pass": n if name == "dec_c_ring" else nr, "mismatches": bad,
                     "checksum": "%016x" % first, "seconds": round(time.time() - t0, 2)}
        print(json.dumps({name: out[name]}), flush=True)
    # the persistent consumer fed vector by vector (8192-sample held pushes, 2^28 samples per pass): the regime in which the
    # host and the resident kernel run neck and neck and every fill path of the kernel is taken.  From the second pass on the
    # stream repeats itself, so every pass must write the same words.
    import ctypes as C
    n2 = 1 << 28
    x.free(); y.free()
    x = ctx.alloc(8 * n2 + 256); y = ctx.alloc(n2 + 8 * 8192 + 256)
    ctx.synth_noise(x, 2 * n2)
    pipe = sdr_b200.pipeFirDecimator(d, 8192)
    L.check(L.lib.sdr_pipe_set_persistent(pipe.h, n2))
    n_out = C.c_longlong()
    t0 = time.time()
    first, bad, p_passes = None, 0, max(50, passes // 5)
    for i in range(p_passes + 1):
        L.check(L.lib.sdr_pipe_run(pipe.h, pipe.h, x.ptr, 8192, n2 // 8192, L.SDR_DEVICE_HELD, y.ptr, n2 // 8 + 8192, L.SDR_DEVICE, C.byref(n_out)))
        if i == 0:
            continue
        cs = ctx.checksum32(y, 2 * n_out.value)
        if first is None:
            first = cs
        elif cs != first:
            bad += 1
    out["dec_c_ring_persist"] = {"kernel": "dec_c_ring_persist<128,8,8,32>", "passes": p_passes, "samples_per_pass": n2, "pushes_per_pass": n2 // 8192,
                                 "mismatches": bad, "checksum": "%016x" % first, "seconds": round(time.time() - t0, 2)}
    print(json.dumps({"dec_c_ring_persist": out["dec_c_ring_persist"]}), flush=True)
    assert all(v["mismatches"] == 0 for v in out.values()), out
    print("SOAK_OK")


if __name__ == "__main__":
    main()
