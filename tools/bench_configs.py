#!/usr/bin/env python
"""Device-resident throughput of the non-headline BASELINE.json configs (measurement tool; prints one JSON line per
config).  cfg1: 64-tap symmetric real filter; cfg3: 3/10 resampler, 90 taps, real; elementwise stages; cfg4 chain."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import sdr_b200  # noqa: E402
import synth  # noqa: E402
from sdr_b200 import _lib as L  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timed(ctx, fn, steps=10, warmup=3):
    for _ in range(warmup):
        fn()
    ctx.sync()
    e0, e1 = sdr_b200.Event(ctx), sdr_b200.Event(ctx)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    return e0.elapsed_ms(e1) / steps


def main():
    log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 27
    n = 1 << log2n
    ctx = sdr_b200.default_context()
    x = ctx.alloc(8 * n + 256)
    y = ctx.alloc(8 * n + 256)
    ctx.synth_noise(x, 2 * n)
    out = []

    def report(name, ms, samples, bytes_per_sample, kernel=""):
        gs = samples / (ms * 1e-3) / 1e9
        out.append({"config": name, "ms": ms, "Gsamples_per_s": gs, "algo_GBps": gs * bytes_per_sample,
                    "hbm_frac": gs * bytes_per_sample / PEAK, "samples": samples, "kernel": kernel})
        print(json.dumps(out[-1]), flush=True)

    # cfg1: fastFilterSymR, 64 taps (32 half taps), real
    half = synth.windowed_sinc_taps(64, 1 / 4)[:32]
    f = sdr_b200.cudaFilterSymR(half, ctx=ctx)
    nr = 2 * n   # real samples available in the buffer
    num = nr - 64 + 1
    ms = timed(ctx, lambda: L.check(L.lib.sdr_filter_stream(f.handle, x.ptr, nr, y.ptr, num)))
    report("cfg1 filter 64-tap sym real", ms, nr, 8.0, f.last_kernel())
    ctx.set_fast_fir(True)
    ms = timed(ctx, lambda: L.check(L.lib.sdr_filter_stream(f.handle, x.ptr, nr, y.ptr, num)))
    report("cfg1 filter 64-tap sym real, 2-parallel fast-FIR arithmetic (opt-in)", ms, nr, 8.0, f.last_kernel())
    ctx.set_fast_fir(False)
    # cfg3: fastResamplerR 3/10, 90 taps, real
    t_res = synth.windowed_sinc_taps(90, 1 / 20, gain=3.0)
    r = sdr_b200.cudaResamplerR(3, 10, t_res, ctx=ctx, sizeMultiple=8)
    num = (nr * 3 - r.numCoeffsR) // 10 + 1
    ms = timed(ctx, lambda: L.check(L.lib.sdr_resample_stream(r.handle, x.ptr, nr, y.ptr, num)))
    report("cfg3 resample 3/10 90-tap real", ms, nr, 5.2)
    # cfg2 for reference
    taps = synth.windowed_sinc_taps(128, 1 / 16)
    d = sdr_b200.cudaDecimatorC(8, taps, ctx=ctx, sizeMultiple=4)
    num = (n - 128) // 8 + 1
    ms = timed(ctx, lambda: L.check(L.lib.sdr_decimate_stream(d.handle, x.ptr, n, y.ptr, num)))
    report("cfg2 decimate-by-8 128-tap complex", ms, n, 9.0, d.last_kernel())
    # element-wise stages (K4, K5, P4), device resident, through the layer-1 kernels' device entry points
    nbytes = 2 * n
    bbuf = ctx.alloc(nbytes)
    ctx.synth_bytes(bbuf, nbytes)
    ms = timed(ctx, lambda: L.check(L.lib.sdr_dev_convert_u8(ctx.h, bbuf.ptr, y.ptr, nbytes)))
    report("K4 convert u8 IQ -> complex float (per IQ pair)", ms, n, 10.0)
    ms = timed(ctx, lambda: L.check(L.lib.sdr_dev_scale(ctx.h, 0.2, x.ptr, y.ptr, 2 * n)))
    report("K5 scale (per float)", ms, 2 * n, 8.0)
    ms = timed(ctx, lambda: L.check(L.lib.sdr_dev_fm_demod(ctx.h, 0.0, 0.0, x.ptr, y.ptr, n)))
    report("P4 fmDemod (per complex sample)", ms, n, 12.0)
    bbuf.free()
    try:
        # dcBlocker (filter.c:152), chunk-parallel by speculation (csrc/dc_spec.cuh): default tuning, then a sweep of the
        # chunk length and of the exact warm-up; the serial kernel on a 2^22-sample piece for scale
        d_fin = ctx.alloc(8)
        for label, tune in (("auto", (0, -1, -1)), ("chunk 1024", (1024, -1, -1)), ("chunk 2048", (2048, -1, -1)),
                            ("chunk 4096", (4096, -1, -1)), ("chunk 8192", (8192, -1, -1)), ("chunk 2048, K2 3072", (2048, -1, 3072)),
                            ("chunk 2048, K1 4096", (2048, 4096, -1))):
            ctx.dc_tuning(tune[0], tune[1], tune[2], -1)
            st0, _ = ctx.dc_stats()
            ms = timed(ctx, lambda: ctx.dc_blocker(x.ptr, y.ptr, 2 * n, d_fin.ptr), steps=5, warmup=2)
            st1, par = ctx.dc_stats()
            report(f"dcBlocker chunk-parallel [{label}] (per float; {st1[1] - st0[1]} chunks, {st1[2] - st0[2]} repaired in 7 calls)",
                   ms, 2 * n, 8.0, "k_dc_spec" if par else "k_dc_blocker")
        ctx.dc_tuning(0, -1, -1, 1 << 40)
        ms = timed(ctx, lambda: ctx.dc_blocker(x.ptr, y.ptr, 1 << 22, d_fin.ptr), steps=2, warmup=1)
        report("dcBlocker serial one-lane kernel (per float)", ms, 1 << 22, 8.0, "k_dc_blocker")
        ctx.dc_tuning()
        d_fin.free()
    except Exception as e:   # measurement aid only
        print(json.dumps({"config": "dcBlocker", "error": str(e)}), flush=True)

    # cfg4: the FM chain through connected device pipes, u8 IQ in
    nb = n   # IQ pairs
    raw = ctx.alloc(2 * nb)
    ctx.synth_bytes(raw, 2 * nb)
    fil = sdr_b200.cudaFilterSymR(half, ctx=ctx)

    # the fused front end alone: u8 IQ -> phase, device resident
    fe = sdr_b200.pipeFmFrontEnd(d, 8192)
    L.check(L.lib.sdr_pipe_set_batch(fe.h, 1 << 23))
    n_out0 = C.c_longlong()

    def run_fe():
        L.check(L.lib.sdr_pipe_run(fe.h, fe.h, raw.ptr, 1 << 27, (2 * nb) >> 27, L.SDR_DEVICE, y.ptr, 2 * n, L.SDR_DEVICE, C.byref(n_out0)))
    ms = timed(ctx, run_fe, steps=5, warmup=2)
    report("cfg4 fused front end alone (u8 IQ -> convert -> decimate-by-8 -> fmDemod)", ms, nb, 2.5, L.lib.sdr_pipe_last_kernel(fe.h).decode())
    fe.close()

    # recorded-IQ replay: fromHandle (Serialize.hs:82) on a tmpfs file -> pinned ring -> fused front end, output discarded
    try:
        path = "/dev/shm/sdr_b200_iq.u8"
        nrep = min(nb, 1 << 28)            # IQ pairs in the file (2 bytes each)
        host = raw.to_host(np.uint8, 2 * nrep)
        host.tofile(path)
        del host
        import time
        for vec_bytes in (16384, 1 << 20):
            fe2 = sdr_b200.pipeFmFrontEnd(d, 8192)
            L.check(L.lib.sdr_pipe_set_batch(fe2.h, 1 << 21))
            with open(path, "rb") as fi:
                t0 = time.perf_counter()
                st = sdr_b200.serialize.runHandles(fe2, fe2, vec_bytes, fi, None)
                dt = time.perf_counter() - t0
            report(f"file replay (tmpfs) -> fromHandle {vec_bytes}-byte vectors -> fused FM front end (wall clock; read() {st.read_seconds:.3f} s)",
                   dt * 1e3, st.elements_in // 2, 2.5, L.lib.sdr_pipe_last_kernel(fe2.h).decode())
            fe2.close()
        os.remove(path)
    except Exception as e:   # measurement aid only
        print(json.dumps({"config": "file replay", "error": str(e)}), flush=True)

    def build_chain(fused):
        if fused:
            head = sdr_b200.pipeFmFrontEnd(d, 8192)
            stages = [head]
        else:
            p0 = sdr_b200.pipeConvertU8(ctx)
            p1 = sdr_b200.pipeFirDecimator(d, 8192)
            p2 = sdr_b200.pipeFmDemod(ctx)
            p0.connect(p1).connect(p2)
            head, stages = p0, [p0, p1, p2]
        p3 = sdr_b200.pipeFirResampler(r, 8192)
        p4 = sdr_b200.pipeFirFilter(fil, 8192)
        p5 = sdr_b200.pipeScale(0.2, ctx)
        stages[-1].connect(p3).connect(p4).connect(p5)
        stages += [p3, p4, p5]
        for p in stages:
            if p is not p5 and p.in_dtype != np.uint8 or (fused and p is head):
                try:
                    L.check(L.lib.sdr_pipe_set_batch(p.h, 1 << 21))
                except sdr_b200.SdrError:
                    pass
        return head, p5, stages

    for fused, chunk in ((False, 1 << 25), (True, 1 << 25), (True, 1 << 26), (True, 1 << 27), (True, 1 << 28)):
        if chunk > 2 * nb:
            continue
        head, tail, stages = build_chain(fused)
        n_out = C.c_longlong()
        # chunk = bytes per push (device memory): 2^25 = 16M IQ pairs

        def run():
            L.check(L.lib.sdr_pipe_run(head.h, tail.h, raw.ptr, chunk, (2 * nb) // chunk, L.SDR_DEVICE, y.ptr, 2 * n, L.SDR_DEVICE,
                                       C.byref(n_out)))
        ms = timed(ctx, run, steps=5, warmup=2)
        kern = L.lib.sdr_pipe_last_kernel(head.h).decode() if fused else ""
        report("cfg4 FM chain u8 IQ -> audio, device pipes, " + ("fused front end" if fused else "stage by stage") + f", {chunk >> 20} MiB pushes", ms, nb, 2.15, kern)
        for p in stages:
            p.close()

    # cfg4 end to end: u8 IQ vectors in pinned HOST memory -> fused chain -> audio vectors in pinned host memory
    # (sdr_pipe_run, 8192-IQ-pair vectors; the link carries 2 B per input sample instead of cfg2's 8)
    try:
        import time
        nh = min(nb, 1 << 27)
        hin = sdr_b200.PinnedArray(np.uint8, 2 * nh)
        L.check(L.lib.sdr_memcpy_d2h(ctx.h, hin.p, raw.ptr, 2 * nh))
        ctx.sync()
        hout = sdr_b200.PinnedArray(np.float32, nh // 16)
        head, tail, stages = build_chain(True)
        n_out = C.c_longlong()
        best = None
        for rep in range(4):
            t0 = time.perf_counter()
            L.check(L.lib.sdr_pipe_run(head.h, tail.h, hin.p, 16384, (2 * nh) // 16384, L.SDR_HOST_PINNED, hout.p, nh // 16,
                                       L.SDR_HOST_PINNED, C.byref(n_out)))
            dt = time.perf_counter() - t0
            best = dt if best is None or dt < best else best
        report(f"cfg4 FM chain END TO END: pinned host u8 IQ vectors (16384 B) -> fused chain -> pinned host audio (wall clock, best of 4; "
               f"{2 * nh / best / 1e9:.1f} GB/s over PCIe, {n_out.value} audio samples out)", best * 1e3, nh, 2.15,
               L.lib.sdr_pipe_last_kernel(head.h).decode())
        for p in stages:
            p.close()
        hin.free(); hout.free()
    except Exception as e:   # measurement aid only
        print(json.dumps({"config": "cfg4 end to end", "error": str(e)}), flush=True)


if __name__ == "__main__":
    main()
