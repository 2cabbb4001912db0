#!/usr/bin/env python
"""cfg3 (3/10 resampler, 90 taps, real) device-resident throughput for the current SDR_B200_RES_S; one JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import sdr_b200  # noqa: E402
import synth  # noqa: E402
from sdr_b200 import _lib as L  # noqa: E402

n = 1 << 28
ctx = sdr_b200.default_context()
x, y = ctx.alloc(4 * n + 256), ctx.alloc(4 * n // 3 + 256)
ctx.synth_noise(x, n)
r = sdr_b200.cudaResamplerR(3, 10, synth.windowed_sinc_taps(90, 1 / 20, gain=3.0), ctx=ctx, sizeMultiple=8)
num = (n * 3 - r.numCoeffsR) // 10 + 1
for _ in range(3):
    L.check(L.lib.sdr_resample_stream(r.handle, x.ptr, n, y.ptr, num))
ctx.sync()
e0, e1 = sdr_b200.Event(ctx), sdr_b200.Event(ctx)
e0.record()
for _ in range(10):
    L.check(L.lib.sdr_resample_stream(r.handle, x.ptr, n, y.ptr, num))
e1.record()
ms = e0.elapsed_ms(e1) / 10
print(json.dumps({"S": os.environ.get("SDR_B200_RES_S", "2"), "kernel": r.last_kernel(), "ms": round(ms, 4),
                  "Gsamples_per_s": round(n / ms / 1e6, 1), "hbm_frac": round(n / ms / 1e6 * 5.2 / 6545.9, 3),
                  "checksum": "%016x" % ctx.checksum32(y, num)}), flush=True)
