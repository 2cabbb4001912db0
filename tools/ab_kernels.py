#!/usr/bin/env python
"""A/B of the same kernel in two builds of the library on the same box (measurement aid): loads each .so with ctypes,
runs the cfg1 / cfg2 / cfg3 stream entry points, alternating, and prints ms per launch."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Lib:
    def __init__(self, path):
        self.l = C.CDLL(path)
        self.l.sdr_last_error.restype = C.c_char_p
        self.ctx = C.c_void_p()
        self.ck(self.l.sdr_ctx_create(0, C.byref(self.ctx)))

    def ck(self, st):
        assert st == 0, self.l.sdr_last_error()

    def alloc(self, n):
        p = C.c_void_p()
        self.ck(self.l.sdr_dev_alloc(self.ctx, C.c_size_t(n), C.byref(p)))
        return p

    def timed(self, fn, steps=10, warm=3):
        e0, e1 = C.c_void_p(), C.c_void_p()
        self.ck(self.l.sdr_event_create(self.ctx, C.byref(e0))); self.ck(self.l.sdr_event_create(self.ctx, C.byref(e1)))
        for _ in range(warm):
            fn()
        self.ck(self.l.sdr_ctx_sync(self.ctx))
        self.ck(self.l.sdr_event_record(self.ctx, e0))
        for _ in range(steps):
            fn()
        self.ck(self.l.sdr_event_record(self.ctx, e1))
        ms = C.c_float()
        self.ck(self.l.sdr_event_elapsed_ms(e0, e1, C.byref(ms)))
        return ms.value / steps


def taps(n, cutoff, gain=1.0):
    k = np.arange(n, dtype=np.float64) - (n - 1) / 2.0
    return (gain * np.sinc(2 * cutoff * k) * 2 * cutoff * (0.54 - 0.46 * np.cos(2 * np.pi * np.arange(n) / (n - 1)))).astype(np.float32)


def main():
    import time
    paths = sys.argv[1:]
    libs = [Lib(p) for p in paths]
    n = 1 << 27
    cases = []
    for lb in libs:
        x, y = lb.alloc(8 * n + 256), lb.alloc(8 * n + 256)
        lb.ck(lb.l.sdr_synth_noise(lb.ctx, x, C.c_longlong(2 * n), C.c_longlong(0), C.c_uint32(1)))
        half = taps(64, 1 / 4)[:32]
        f = C.c_void_p()
        lb.ck(lb.l.sdr_filter_create_sym(lb.ctx, 0, half.ctypes.data_as(C.c_void_p), 32, C.byref(f)))
        t128 = taps(128, 1 / 16)
        d = C.c_void_p()
        lb.ck(lb.l.sdr_decimator_create(lb.ctx, 1, 8, t128.ctypes.data_as(C.c_void_p), 128, 4, C.byref(d)))
        t90 = taps(90, 1 / 20, 3.0)
        r = C.c_void_p()
        lb.ck(lb.l.sdr_resampler_create(lb.ctx, 0, 3, 10, t90.ctypes.data_as(C.c_void_p), 90, 8, C.byref(r)))
        nr = 2 * n
        cases.append({
            "cfg1": lambda lb=lb, f=f, x=x, y=y: lb.ck(lb.l.sdr_filter_stream(f, x, C.c_longlong(nr), y, C.c_longlong(nr - 63))),
            "cfg2": lambda lb=lb, d=d, x=x, y=y: lb.ck(lb.l.sdr_decimate_stream(d, x, C.c_longlong(n), y, C.c_longlong((n - 128) // 8 + 1))),
            "cfg3": lambda lb=lb, r=r, x=x, y=y: lb.ck(lb.l.sdr_resample_stream(r, x, C.c_longlong(nr), y, C.c_longlong((nr * 3 - 96) // 10 + 1))),
        })
    for rep in range(3):
        for name in ("cfg1", "cfg2", "cfg3"):
            row = []
            for lb, cs in zip(libs, cases):
                time.sleep(0.3)   # cool down: every measurement starts in the burst regime
                row.append(round(lb.timed(cs[name]), 4))
            print(name, "ms per launch:", dict(zip([os.path.basename(p) for p in paths], row)), flush=True)


if __name__ == "__main__":
    main()
