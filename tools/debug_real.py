import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import sdr_b200, synth
from sdr_b200 import _lib as L
ctx = sdr_b200.default_context()
taps = synth.windowed_sinc_taps(90, 1 / 20, gain=3.0)
r = sdr_b200.cudaResamplerR(3, 10, taps, sizeMultiple=8)
for log2n in (25, 26, 27):
    n = 1 << log2n
    num = (n * 3 - r.numCoeffsR) // 10 + 1
    x = ctx.alloc(4 * n + 64); y = ctx.alloc(4 * num + 64); y2 = ctx.alloc(4 * num + 64)
    ctx.synth_noise(x, n)
    L.check(L.lib.sdr_resample_stream(r.handle, x.ptr, n, y.ptr, num)); ctx.sync()
    k1 = r.last_kernel()
    ctx.synth_noise(x, n, first_float=0, offset_bytes=4)
    L.check(L.lib.sdr_resample_stream(r.handle, x.at(4), n, y2.ptr, num)); ctx.sync()
    a = y.to_host(np.float32, num); b = y2.to_host(np.float32, num)
    bad = np.nonzero(a != b)[0]
    print(log2n, k1, r.last_kernel(), "mismatches", len(bad), bad[:10], bad[-5:] if len(bad) else "")
    if len(bad):
        # which one is right?  float64 model on a window around the first mismatch
        k0 = int(bad[0]) // 3 * 3
        xs = synth.noise(4000, first=(k0 * 10 + 2) // 3).astype(np.float64)
        want = []
        for k in range(k0, k0 + 30):
            f = (-k * 10) % 3; i = -((-k * 10) // 3) - (k0 * 10 + 2) // 3
            tp = taps[f::3].astype(np.float64)
            want.append(float(np.dot(tp, xs[i:i + len(tp)])))
        want = np.array(want)
        print(" tuned  err", np.abs(a[k0:k0 + 30] - want).max(), " generic err", np.abs(b[k0:k0 + 30] - want).max())
        d = np.diff(bad); print(" slot of first bad:", bad[0] / 1152.0, "runs:", np.unique(d)[:10])
    for q in (x, y, y2): q.free()
if os.environ.get("SKIP28"): sys.exit(0)
print("fir 2^28")
half = synth.windowed_sinc_taps(64, 1 / 4)[:32]
f = sdr_b200.cudaFilterSymR(half)
n = 1 << 28
x = ctx.alloc(4 * n + 256); y = ctx.alloc(4 * n + 256)
ctx.synth_noise(x, n); ctx.sync()
L.check(L.lib.sdr_filter_stream(f.handle, x.ptr, n, y.ptr, n - 63)); ctx.sync(); print("fir ok", f.last_kernel())
L.check(L.lib.sdr_resample_stream(r.handle, x.ptr, n, y.ptr, (n * 3 - 96) // 10 + 1)); ctx.sync(); print("res ok", r.last_kernel())
