#!/usr/bin/env python
"""headline decimator with a reduced grid (SDR_B200_GRID): does a power-of-two byte stride between the CTAs' streams hurt? (measurement aid)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sdr_b200
from sdr_b200 import _lib as L
ctx = sdr_b200.default_context()
n = 1 << 28
x = ctx.alloc(8 * n + 256); y = ctx.alloc(n + 256)
ctx.synth_noise(x, 2 * n)
d = sdr_b200.cudaDecimatorC(8, sdr_b200.windowed_sinc_taps(128, 1 / 16), ctx=ctx, sizeMultiple=4)
num = (n - 128) // 8 + 1
for rep in range(2):
    time.sleep(0.3)
    for _ in range(2):
        L.check(L.lib.sdr_decimate_stream(d.handle, x.ptr, n, y.ptr, num))
    ctx.sync()
    e0, e1 = sdr_b200.Event(ctx), sdr_b200.Event(ctx)
    e0.record()
    for _ in range(4):
        L.check(L.lib.sdr_decimate_stream(d.handle, x.ptr, n, y.ptr, num))
    e1.record()
    ms = e0.elapsed_ms(e1) / 4
print(d.last_kernel(), "grid", os.environ.get("SDR_B200_GRID", "148"), round(ms, 4), "ms", round(n / ms / 1e6, 1), "Gs/s", flush=True)
