"""PCIe probe behind bench.py's e2e number (measurement tool): sdr_pipe_run from pinned host vectors at several batch
sizes, next to plain / chunked / interleaved cudaMemcpy rates on the same box."""
import sys, os, ctypes as C, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, sdr_b200, synth
from sdr_b200 import _lib as L
ctx = sdr_b200.default_context()
BUF = 8192; n_vecs = 1 << 14   # 2^27 samples = 1 GiB
taps = synth.windowed_sinc_taps(128, 1 / 16)
dec = sdr_b200.cudaDecimatorC(8, taps, ctx=ctx, sizeMultiple=4)
hin = sdr_b200.PinnedArray(np.float32, 2 * n_vecs * BUF); hin.array[:] = 1.0
out_cap = (n_vecs * BUF // 8 // BUF + 1) * BUF
hout = sdr_b200.PinnedArray(np.float32, 2 * out_cap)
n_out = C.c_longlong()
for batch in (64, 512, 4096):
    pipe = sdr_b200.pipeFirDecimator(dec, BUF)
    L.check(L.lib.sdr_pipe_set_batch(pipe.h, batch * BUF))
    for rep in range(4):
        t0 = time.perf_counter()
        L.check(L.lib.sdr_pipe_run(pipe.h, pipe.h, hin.p, BUF, n_vecs, L.SDR_HOST_PINNED, hout.p, out_cap, L.SDR_HOST_PINNED, C.byref(n_out)))
        dt = time.perf_counter() - t0
        print("batch", batch, "rep", rep, "%.2f ms" % (dt * 1e3), "%.2f Gs/s" % (n_vecs * BUF / dt / 1e9), "H2D %.1f GB/s" % (8 * n_vecs * BUF / dt / 1e9), n_out.value, flush=True)
    pipe.close()
# raw: H2D only, D2H only
d = ctx.alloc(hin.array.nbytes)
for rep in range(3):
    t0 = time.perf_counter(); L.check(L.lib.sdr_memcpy_h2d(ctx.h, d.ptr, hin.p, hin.array.nbytes)); ctx.sync(); dt = time.perf_counter() - t0
    print("raw H2D %.1f GB/s" % (hin.array.nbytes / dt / 1e9))
for rep in range(3):
    t0 = time.perf_counter(); L.check(L.lib.sdr_memcpy_d2h(ctx.h, hout.p, d.ptr, hout.array.nbytes)); ctx.sync(); dt = time.perf_counter() - t0
    print("raw D2H %.1f GB/s" % (hout.array.nbytes / dt / 1e9))
# chunked H2D 8 MB pieces
t0 = time.perf_counter()
step = 8 << 20
for off in range(0, hin.array.nbytes, step):
    L.check(L.lib.sdr_memcpy_h2d(ctx.h, C.c_void_p(d.ptr.value + off), C.c_void_p(hin.p.value + off), step))
ctx.sync(); dt = time.perf_counter() - t0
print("chunked 8MB H2D %.1f GB/s" % (hin.array.nbytes / dt / 1e9))
for dst_off in (0, 960, 4096, 65536 + 960, 1 << 21):
    d2 = ctx.alloc(hin.array.nbytes + (4 << 20))
    for rep in range(2):
        t0 = time.perf_counter()
        for off in range(0, hin.array.nbytes, step):
            L.check(L.lib.sdr_memcpy_h2d(ctx.h, C.c_void_p(d2.ptr.value + dst_off + off), C.c_void_p(hin.p.value + off), step))
        ctx.sync(); dt = time.perf_counter() - t0
    print("chunked 8MB H2D dst offset", dst_off, "%.1f GB/s" % (hin.array.nbytes / dt / 1e9))
    d2.free()
d2 = ctx.alloc(hin.array.nbytes)
for every in (4, 16, 64):
    for rep in range(2):
        t0 = time.perf_counter(); k = 0; o = 0
        for off in range(0, hin.array.nbytes, step):
            L.check(L.lib.sdr_memcpy_h2d(ctx.h, C.c_void_p(d2.ptr.value + off), C.c_void_p(hin.p.value + off), step))
            k += 1
            if k % every == 0:
                nb = every * (1 << 20)
                L.check(L.lib.sdr_memcpy_d2h(ctx.h, C.c_void_p(hout.p.value + o), C.c_void_p(d2.ptr.value + o), nb)); o += nb
        ctx.sync(); dt = time.perf_counter() - t0
    print("H2D 8MB chunks with a D2H of 1/8 the bytes every", every, "chunks: H2D rate %.1f GB/s" % (hin.array.nbytes / dt / 1e9))
