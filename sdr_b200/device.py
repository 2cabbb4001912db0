"""Device context, device / pinned buffers and CUDA-event timing over the C ABI (no torch types)."""
import ctypes as C

import numpy as np

from . import _lib as L


class Context:
    """sdr_ctx_t: one device + one stream.  Plays the role CPUInfo/getCPUInfo play in the reference's dispatch
    (hs_sources/SDR/CPUID.hs:61-75)."""

    def __init__(self, device=0, arith=L.SDR_ARITH_FAST):
        h = C.c_void_p()
        L.check(L.lib.sdr_ctx_create(device, C.byref(h)))
        self.h = h
        self.device = device
        if arith != L.SDR_ARITH_FAST:
            self.set_arith(arith)

    def set_arith(self, mode):
        L.check(L.lib.sdr_ctx_set_arith(self.h, mode))

    def set_fast_fir(self, on=True):
        """real 32/64/128-tap stride-1 filters: 2-parallel fast-FIR arithmetic in the tuned kernel (csrc/fir_ffa.cuh)"""
        L.check(L.lib.sdr_ctx_set_fast_fir(self.h, int(bool(on))))

    def sync(self):
        L.check(L.lib.sdr_ctx_sync(self.h))

    @property
    def sm_count(self):
        n = C.c_int()
        L.check(L.lib.sdr_ctx_sm_count(self.h, C.byref(n)))
        return n.value

    @property
    def launches(self):
        n = C.c_longlong()
        L.check(L.lib.sdr_ctx_launch_count(self.h, C.byref(n)))
        return n.value

    def alloc(self, nbytes):
        return DeviceBuffer(self, nbytes)

    def to_device(self, arr):
        arr = np.ascontiguousarray(arr)
        buf = DeviceBuffer(self, arr.nbytes)
        L.check(L.lib.sdr_memcpy_h2d(self.h, buf.ptr, L.ptr(arr), arr.nbytes))
        self.sync()
        return buf

    def flush_l2(self):
        L.check(L.lib.sdr_flush_l2(self.h))

    def checksum32(self, buf, n_words, first_word=0, offset_bytes=0):
        s = C.c_uint64()
        L.check(L.lib.sdr_checksum32(self.h, C.c_void_p(buf.ptr.value + offset_bytes), n_words, first_word, C.byref(s)))
        return s.value

    def synth_noise(self, buf, n_floats, first_float=0, seed=0x5D2B200, offset_bytes=0):
        L.check(L.lib.sdr_synth_noise(self.h, C.c_void_p(buf.ptr.value + offset_bytes), n_floats, first_float, seed))

    def synth_bytes(self, buf, n_bytes, first_byte=0, seed=0x5D2B200, offset_bytes=0):
        L.check(L.lib.sdr_synth_bytes(self.h, C.c_void_p(buf.ptr.value + offset_bytes), n_bytes, first_byte, seed))

    def dc_blocker(self, d_in, d_out, n, d_final2, last_sample=0.0, last_output=0.0):
        """dcBlocker (filter.c:152-161) on device buffers (ctypes pointers); enqueue-only"""
        L.check(L.lib.sdr_dev_dc_blocker(self.h, last_sample, last_output, d_in, d_out, n, d_final2))

    def dc_tuning(self, chunk=0, cheap_warmup=-1, exact_warmup=-1, min_parallel=-1):
        """speculation parameters of the chunk-parallel dcBlocker (csrc/dc_spec.cuh); the defaults restore automatic"""
        L.check(L.lib.sdr_dc_blocker_tuning(self.h, chunk, cheap_warmup, exact_warmup, min_parallel))

    def dc_stats(self):
        """(parallel calls, chunks, chunks repaired, samples rewritten by repairs), last call took the parallel path"""
        st = (C.c_longlong * 4)()
        par = C.c_int()
        L.check(L.lib.sdr_dc_blocker_stats(self.h, st, C.byref(par)))
        return tuple(int(v) for v in st), bool(par.value)

    def close(self):
        if self.h:
            L.lib.sdr_ctx_destroy(self.h)
            self.h = None


class DeviceBuffer:
    def __init__(self, ctx, nbytes):
        self.ctx = ctx
        self.nbytes = int(nbytes)
        p = C.c_void_p()
        L.check(L.lib.sdr_dev_alloc(ctx.h, self.nbytes, C.byref(p)))
        self.ptr = p

    def at(self, offset_bytes):
        return C.c_void_p(self.ptr.value + int(offset_bytes))

    def to_host(self, dtype, count, offset_bytes=0):
        out = np.empty(count, dtype)
        L.check(L.lib.sdr_memcpy_d2h(self.ctx.h, L.ptr(out), self.at(offset_bytes), out.nbytes))
        self.ctx.sync()
        return out

    def upload(self, arr, offset_bytes=0):
        arr = np.ascontiguousarray(arr)
        L.check(L.lib.sdr_memcpy_h2d(self.ctx.h, self.at(offset_bytes), L.ptr(arr), arr.nbytes))
        self.ctx.sync()

    def free(self):
        if self.ptr:
            L.lib.sdr_dev_free(self.ctx.h, self.ptr)
            self.ptr = None


class PinnedArray:
    """page-locked host memory viewed as a numpy array"""

    def __init__(self, dtype, count):
        self.dtype = np.dtype(dtype)
        self.count = int(count)
        p = C.c_void_p()
        L.check(L.lib.sdr_host_alloc_pinned(self.count * self.dtype.itemsize, C.byref(p)))
        self.p = p
        buf = (C.c_char * (self.count * self.dtype.itemsize)).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=self.count)

    def free(self):
        if self.p:
            self.array = None
            L.lib.sdr_host_free_pinned(self.p)
            self.p = None


class Event:
    def __init__(self, ctx):
        self.ctx = ctx
        h = C.c_void_p()
        L.check(L.lib.sdr_event_create(ctx.h, C.byref(h)))
        self.h = h

    def record(self):
        L.check(L.lib.sdr_event_record(self.ctx.h, self.h))

    def elapsed_ms(self, stop):
        ms = C.c_float()
        L.check(L.lib.sdr_event_elapsed_ms(self.h, stop.h, C.byref(ms)))
        return ms.value

    def destroy(self):
        if self.h:
            L.lib.sdr_event_destroy(self.h)
            self.h = None


def has_cuda():
    """`hasCUDA`: the predicate a featureSelect entry would test (CPUID.hs:100-104)."""
    return bool(L.lib.sdr_has_cuda())


def featureSelect(info, default, options):
    """``featureSelect :: CPUInfo -> a -> [(CPUInfo -> Bool, a)] -> a`` (hs_sources/SDR/CPUID.hs:100-104): the first
    implementation whose predicate accepts `info`, else `default`.  `info` is whatever the predicates test -- for this
    backend ``hasCUDA`` below ignores it and asks the library; a caller mixing it with its own predicates passes its own
    info object through unchanged, exactly as the reference threads its CPUInfo."""
    for pred, impl in options:
        if pred(info):
            return impl
    return default


def hasCUDA(_info=None):
    """the predicate that goes in front of hasAVX / hasSSE42 in a featureSelect list (CPUID.hs:87-104)"""
    return has_cuda()


def device_count():
    n = C.c_int()
    L.check(L.lib.sdr_device_count(C.byref(n)))
    return n.value
