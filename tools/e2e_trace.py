import sys, os, ctypes as C, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, sdr_b200, synth
from sdr_b200 import _lib as L
ctx = sdr_b200.default_context()
BUF = 8192; n_vecs = 1 << 13
dec = sdr_b200.cudaDecimatorC(8, synth.windowed_sinc_taps(128, 1 / 16), ctx=ctx, sizeMultiple=4)
hin = sdr_b200.PinnedArray(np.float32, 2 * n_vecs * BUF); hin.array[:] = 1.0
out_cap = (n_vecs * BUF // 8 // BUF + 1) * BUF
hout = sdr_b200.PinnedArray(np.float32, 2 * out_cap)
n_out = C.c_longlong()
pipe = sdr_b200.pipeFirDecimator(dec, BUF)
L.check(L.lib.sdr_pipe_set_batch(pipe.h, 2048 * BUF))
for rep in range(3):
    print("=== rep", rep, file=sys.stderr, flush=True)
    t0 = time.perf_counter()
    L.check(L.lib.sdr_pipe_run(pipe.h, pipe.h, hin.p, BUF, n_vecs, L.SDR_HOST_PINNED, hout.p, out_cap, L.SDR_HOST_PINNED, C.byref(n_out)))
    print("rep", rep, "%.2f ms" % ((time.perf_counter() - t0) * 1e3), file=sys.stderr, flush=True)
