#!/usr/bin/env python
"""cfg4 chain (two fused stages, 32 MiB in-place pushes) under SDR_B200_TRACE=2: device-side duration and start time of
every stage step, host issue time (measurement aid).  usage: SDR_B200_TRACE=2 python tools/chain_probe.py [log2n]"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sdr_b200
from sdr_b200 import _lib as L
ctx = sdr_b200.default_context()
log2 = int(sys.argv[1]) if len(sys.argv) > 1 else 27
n = 1 << log2
BUF = 8192
taps = sdr_b200.windowed_sinc_taps(128, 1 / 16)
dec = sdr_b200.cudaDecimatorC(8, taps, sizeMultiple=4, ctx=ctx)
r = sdr_b200.cudaResamplerR(3, 10, sdr_b200.windowed_sinc_taps(90, 1 / 20, gain=3.0), sizeMultiple=8, ctx=ctx)
fil = sdr_b200.cudaFilterSymR(sdr_b200.windowed_sinc_taps(64, 1 / 4)[:32], ctx=ctx)
nbytes = 2 * n
bbuf = ctx.alloc(nbytes + 256)
y = ctx.alloc(8 * n // 8 + 256)
ctx.synth_bytes(bbuf, nbytes)
push = 1 << 25
stages = [sdr_b200.pipeFmFrontEnd(dec, BUF), sdr_b200.pipeFmLowRate(r, BUF, fil, BUF, 0.2)]
stages[0].connect(stages[1])
for p in stages:
    L.check(L.lib.sdr_pipe_set_batch(p.h, 1 << 21))
n_out = C.c_longlong()
def chain():
    L.check(L.lib.sdr_pipe_run(stages[0].h, stages[-1].h, bbuf.ptr, push, nbytes // push, L.SDR_DEVICE_HELD, y.ptr, n, L.SDR_DEVICE, C.byref(n_out)))
for _ in range(2):
    chain()
ctx.sync()
for rep in range(3):
    e0, e1 = sdr_b200.Event(ctx), sdr_b200.Event(ctx)
    t0 = time.perf_counter()
    e0.record()
    chain()
    e1.record()
    t1 = time.perf_counter()
    ms = e0.elapsed_ms(e1)
    print(f"chain pass: device {ms*1e3:.1f} us, host issue {(t1-t0)*1e6:.1f} us, {n/ms/1e6:.1f} Gs/s, out {n_out.value}", flush=True)
for st in stages:
    st.sync()
