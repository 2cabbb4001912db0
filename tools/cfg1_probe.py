#!/usr/bin/env python
"""cfg1 (64-tap symmetric real filter) and the 32 / 128-tap real filters, device resident (measurement aid)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sdr_b200
from sdr_b200 import _lib as L
ctx = sdr_b200.default_context()
n = 1 << 28
x = ctx.alloc(4 * n + 256); y = ctx.alloc(4 * n + 256)
ctx.synth_noise(x, n)
for T in (64, 32, 128):
    f = sdr_b200.cudaFilterSymR(sdr_b200.windowed_sinc_taps(T, 1 / 4)[:T // 2], ctx=ctx)
    for rep in range(3):
        time.sleep(0.3)
        for _ in range(3):
            L.check(L.lib.sdr_filter_stream(f.handle, x.ptr, n, y.ptr, n - T + 1))
        ctx.sync()
        e0, e1 = sdr_b200.Event(ctx), sdr_b200.Event(ctx)
        e0.record()
        for _ in range(8):
            L.check(L.lib.sdr_filter_stream(f.handle, x.ptr, n, y.ptr, n - T + 1))
        e1.record()
        ms = e0.elapsed_ms(e1) / 8
        print("T", T, f.last_kernel(), "PAIR", os.environ.get("SDR_B200_FIR_PAIR", "1"), round(ms, 4), "ms", round(n / ms / 1e6, 1), "Gs/s", "checksum %016x" % ctx.checksum32(y, n - T + 1), flush=True)
