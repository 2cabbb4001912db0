import sys, os, ctypes as C, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, sdr_b200, synth
from sdr_b200 import _lib as L
ctx = sdr_b200.default_context()
n = 1 << 26
raw = ctx.alloc(2 * n); ctx.synth_bytes(raw, 2 * n)
y = ctx.alloc(8 * n)
d = sdr_b200.cudaDecimatorC(8, synth.windowed_sinc_taps(128, 1 / 16), ctx=ctx, sizeMultiple=4)
r = sdr_b200.cudaResamplerR(3, 10, synth.windowed_sinc_taps(90, 1 / 20, gain=3.0), ctx=ctx, sizeMultiple=8)
fil = sdr_b200.cudaFilterSymR(synth.windowed_sinc_taps(64, 1 / 4)[:32], ctx=ctx)
head = sdr_b200.pipeFmFrontEnd(d, 8192)
p3 = sdr_b200.pipeFirResampler(r, 8192); p4 = sdr_b200.pipeFirFilter(fil, 8192); p5 = sdr_b200.pipeScale(0.2, ctx)
head.connect(p3).connect(p4).connect(p5)
for p in (head, p3, p4):
    L.check(L.lib.sdr_pipe_set_batch(p.h, 1 << 21))
n_out = C.c_longlong(); chunk = 1 << 25
def run():
    L.check(L.lib.sdr_pipe_run(head.h, p5.h, raw.ptr, chunk, (2 * n) // chunk, L.SDR_DEVICE, y.ptr, 2 * n, L.SDR_DEVICE, C.byref(n_out)))
for i in range(4):
    t0 = time.perf_counter(); run(); ctx.sync(); print("pass", i, (time.perf_counter() - t0) * 1e3, "ms", n_out.value, ctx.launches)
