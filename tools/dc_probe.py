#!/usr/bin/env python
"""One dcBlocker call (chunk-parallel) and one fmDemod call on 2^27-element device-resident streams: the target of the
ncu captures under profiles/ (ncu -k regex:k_dc_spec / k_fm_demod4)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sdr_b200  # noqa: E402
from sdr_b200 import _lib as L  # noqa: E402

n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 27)
ctx = sdr_b200.default_context()
x, y, fin = ctx.alloc(4 * n), ctx.alloc(4 * n), ctx.alloc(8)
ctx.synth_noise(x, n)
for _ in range(2):
    ctx.dc_blocker(x.ptr, y.ptr, n, fin.ptr)
    L.check(L.lib.sdr_dev_fm_demod(ctx.h, 0.0, 0.0, x.ptr, y.ptr, n // 2))
ctx.sync()
print("dc stats", ctx.dc_stats())
