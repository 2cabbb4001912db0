#!/usr/bin/env python
"""probe of the persistent-consumer bench row (debugging aid)"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sdr_b200
from sdr_b200 import _lib as L
import bench
log2 = int(sys.argv[1]) if len(sys.argv) > 1 else 26
n = 1 << log2
BUF = 8192
ctx = sdr_b200.default_context()
dec = sdr_b200.cudaDecimatorC(8, bench.design_taps(), ctx=ctx, sizeMultiple=4)
x = ctx.alloc(8 * n + 256); y = ctx.alloc(n + 8 * BUF + 256)
ctx.synth_noise(x, 2 * n)
n_out = C.c_longlong()
pipe = sdr_b200.pipeFirDecimator(dec, BUF)
L.check(L.lib.sdr_pipe_set_persistent(pipe.h, n))
for i in range(5):
    t0 = time.perf_counter()
    L.check(L.lib.sdr_pipe_run(pipe.h, pipe.h, x.ptr, BUF, n // BUF, L.SDR_DEVICE_HELD, y.ptr, n // 8 + BUF, L.SDR_DEVICE, C.byref(n_out)))
    ctx.sync()
    print("pass", i, round((time.perf_counter() - t0) * 1e3, 3), "ms", n_out.value, dec.last_kernel(), flush=True)
# everything published at once: one big held vector (pure kernel rate, no catching up with the host)
for i in range(4):
    t0 = time.perf_counter()
    L.check(L.lib.sdr_pipe_push(pipe.h, x.ptr, n, L.SDR_DEVICE_HELD))
    L.check(L.lib.sdr_pipe_sync(pipe.h))
    dt = (time.perf_counter() - t0) * 1e3
    got = 0
    while pipe.ready():
        pipe_out = pipe.pop(); got += len(pipe_out)
    print("one-vector pass", i, round(dt, 3), "ms", got, flush=True)
