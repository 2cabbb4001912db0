"""Host-side mirror of the reference's ``SDR.Filter`` for the CUDA backend.

Same names, argument meaning and error behaviour as ``hs_sources/SDR/Filter.hs``:

* records ``Filter`` / ``Decimator`` / ``Resampler`` (Filter.hs:116-144) whose closures run on the B200 through
  the C ABI (``include/sdr_b200.h`` layer 2);
* constructors ``cudaFilterR/C/SymR``, ``cudaDecimatorR/C/SymR``, ``cudaResamplerR/C`` -- the CUDA members of the
  ``fast*`` families (Filter.hs:193-196, 229-232, 258-261, 311-315, 352-356, 385-389, 468-473, 497-502);
* the streaming stages ``firFilter`` / ``firDecimator`` / ``firResampler`` (Filter.hs:532-727) as generators:
  they pull input vectors from an iterable (the Pipe's ``await``) and yield output vectors of exactly
  ``blockSizeOut`` elements (its ``yield``).  On the device the stream is contiguous, so there is no crossover
  state; ``firFilterRecord`` etc. run the reference's own simple/crossover state machine over the record closures
  for parity checks of the closures themselves.
"""
import ctypes as C
from dataclasses import dataclass
from typing import Any, Callable, Iterable, Iterator

import numpy as np

from . import _lib as L
from .device import Context

_default_ctx = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


def _dtype(cplx):
    return np.complex64 if cplx else np.float32


@dataclass
class Filter:
    """data Filter (Filter.hs:116-120)"""
    numCoeffsF: int
    filterOne: Callable      # count -> bufIn -> out[count]
    filterCross: Callable    # count -> bufLast -> bufNext -> out[count]
    cplx: bool
    handle: Any
    ctx: Context

    def last_kernel(self):
        return L.lib.sdr_filter_last_kernel(self.handle).decode()

    def __del__(self):
        if L is not None and self.handle:
            L.lib.sdr_filter_destroy(self.handle)
            self.handle = None


@dataclass
class Decimator:
    """data Decimator (Filter.hs:126-131)"""
    numCoeffsD: int
    decimationD: int
    decimateOne: Callable
    decimateCross: Callable
    cplx: bool
    handle: Any
    ctx: Context

    def last_kernel(self):
        return L.lib.sdr_decimator_last_kernel(self.handle).decode()

    def __del__(self):
        if L is not None and self.handle:
            L.lib.sdr_decimator_destroy(self.handle)
            self.handle = None


@dataclass
class Resampler:
    """data Resampler (Filter.hs:137-144); the existential state ``dat`` is the pair (group, offset) (:424)"""
    numCoeffsR: int
    decimationR: int
    interpolationR: int
    startDat: Any
    resampleOne: Callable    # dat -> count -> bufIn -> (out[count], (dat', endOffset))
    resampleCross: Callable  # dat -> count -> bufLast -> bufNext -> (out[count], (dat', endOffset))
    cplx: bool
    handle: Any
    ctx: Context

    def last_kernel(self):
        return L.lib.sdr_resampler_last_kernel(self.handle).decode()

    def __del__(self):
        if L is not None and self.handle:
            L.lib.sdr_resampler_destroy(self.handle)
            self.handle = None


def _one(fn, handle, cplx):
    def run(count, buf):
        x = L.as_floats(buf)
        out = np.empty(count, _dtype(cplx))
        L.check(fn(handle, count, L.ptr(x), L.ptr(out), L.SDR_HOST))
        return out
    return run


def _cross(fn, handle, cplx):
    def run(count, last, nxt):
        a, b = L.as_floats(last), L.as_floats(nxt)
        div = 2 if cplx else 1
        out = np.empty(count, _dtype(cplx))
        L.check(fn(handle, count, L.ptr(a), len(a) // div, L.ptr(b), len(b) // div, L.ptr(out), L.SDR_HOST))
        return out
    return run


def _mk_filter(ctx, cplx, coeffs, size_multiple, sym):
    ctx = ctx or default_context()
    c = L.f32(coeffs)
    h = C.c_void_p()
    if sym:
        L.check(L.lib.sdr_filter_create_sym(ctx.h, int(cplx), L.ptr(c), len(c), C.byref(h)))
    else:
        L.check(L.lib.sdr_filter_create(ctx.h, int(cplx), L.ptr(c), len(c), size_multiple, C.byref(h)))
    return Filter(L.lib.sdr_filter_num_coeffs(h), _one(L.lib.sdr_filter_one, h, cplx),
                  _cross(L.lib.sdr_filter_cross, h, cplx), cplx, h, ctx)


def cudaFilterR(coeffs, ctx=None, sizeMultiple=1) -> Filter:
    """fastFilterR (Filter.hs:193-196) on the B200: real data, real taps"""
    return _mk_filter(ctx, False, coeffs, sizeMultiple, False)


def cudaFilterC(coeffs, ctx=None, sizeMultiple=1) -> Filter:
    """fastFilterC (Filter.hs:229-232): complex data, real taps"""
    return _mk_filter(ctx, True, coeffs, sizeMultiple, False)


def cudaFilterSymR(halfCoeffs, ctx=None) -> Filter:
    """fastFilterSymR (Filter.hs:258-261): pass the FIRST HALF of a symmetric tap list"""
    return _mk_filter(ctx, False, halfCoeffs, 1, True)


def _mk_decimator(ctx, cplx, factor, coeffs, size_multiple, sym):
    ctx = ctx or default_context()
    c = L.f32(coeffs)
    h = C.c_void_p()
    if sym:
        L.check(L.lib.sdr_decimator_create_sym(ctx.h, int(cplx), factor, L.ptr(c), len(c), C.byref(h)))
    else:
        L.check(L.lib.sdr_decimator_create(ctx.h, int(cplx), factor, L.ptr(c), len(c), size_multiple, C.byref(h)))
    return Decimator(L.lib.sdr_decimator_num_coeffs(h), L.lib.sdr_decimator_factor(h),
                     _one(L.lib.sdr_decimate_one, h, cplx), _cross(L.lib.sdr_decimate_cross, h, cplx), cplx, h, ctx)


def cudaDecimatorR(factor, coeffs, ctx=None, sizeMultiple=1) -> Decimator:
    """fastDecimatorR (Filter.hs:311-315)"""
    return _mk_decimator(ctx, False, factor, coeffs, sizeMultiple, False)


def cudaDecimatorC(factor, coeffs, ctx=None, sizeMultiple=1) -> Decimator:
    """fastDecimatorC (Filter.hs:352-356): complex data, real taps -- the headline path"""
    return _mk_decimator(ctx, True, factor, coeffs, sizeMultiple, False)


def cudaDecimatorSymR(factor, halfCoeffs, ctx=None) -> Decimator:
    """fastDecimatorSymR (Filter.hs:385-389)"""
    return _mk_decimator(ctx, False, factor, halfCoeffs, 1, True)


def _mk_resampler(ctx, cplx, interpolation, decimation, coeffs, size_multiple):
    ctx = ctx or default_context()
    c = L.f32(coeffs)
    h = C.c_void_p()
    L.check(L.lib.sdr_resampler_create(ctx.h, int(cplx), interpolation, decimation, L.ptr(c), len(c), size_multiple,
                                       C.byref(h)))

    def resample_one(dat, count, buf):
        x = L.as_floats(buf)
        out = np.empty(count, _dtype(cplx))
        d = L.ResamplerDat(dat[0], dat[1])
        end = C.c_int()
        L.check(L.lib.sdr_resample_one(h, C.byref(d), count, L.ptr(x), L.ptr(out), L.SDR_HOST, C.byref(end)))
        return out, ((d.group, d.offset), end.value)

    def resample_cross(dat, count, last, nxt):
        a, b = L.as_floats(last), L.as_floats(nxt)
        div = 2 if cplx else 1
        out = np.empty(count, _dtype(cplx))
        d = L.ResamplerDat(dat[0], dat[1])
        end = C.c_int()
        L.check(L.lib.sdr_resample_cross(h, C.byref(d), count, L.ptr(a), len(a) // div, L.ptr(b), len(b) // div,
                                         L.ptr(out), L.SDR_HOST, C.byref(end)))
        return out, ((d.group, d.offset), end.value)

    return Resampler(L.lib.sdr_resampler_num_coeffs(h), decimation, interpolation, (0, 0), resample_one,
                     resample_cross, cplx, h, ctx)


def cudaResamplerR(interpolation, decimation, coeffs, ctx=None, sizeMultiple=1) -> Resampler:
    """fastResamplerR (Filter.hs:468-473)"""
    return _mk_resampler(ctx, False, interpolation, decimation, coeffs, sizeMultiple)


def cudaResamplerC(interpolation, decimation, coeffs, ctx=None, sizeMultiple=1) -> Resampler:
    """fastResamplerC (Filter.hs:497-502)"""
    return _mk_resampler(ctx, True, interpolation, decimation, coeffs, sizeMultiple)


# -----------------------------------------------------------------------------------------------------------------
# streaming stages over the native pipes (include/sdr_b200.h layer 3)
# -----------------------------------------------------------------------------------------------------------------

class NativePipe:
    """sdr_pipe_t: push = the Pipe's await, pop = its yield."""

    def __init__(self, handle, ctx, in_dtype, out_dtype, owner=None):
        self.h, self.ctx, self.in_dtype, self.out_dtype, self.owner = handle, ctx, np.dtype(in_dtype), np.dtype(out_dtype), owner
        self._down = None

    def push(self, vec):
        v = np.ascontiguousarray(vec, dtype=self.in_dtype)
        L.check(L.lib.sdr_pipe_push(self.h, L.ptr(v), len(v), L.SDR_HOST))

    def push_device(self, dptr, n, held=False):
        """push a vector that already lives in device memory (n input elements at `dptr`).  held=True (SDR_DEVICE_HELD): the
        stage reads it in place -- the caller leaves it untouched until sync()"""
        L.check(L.lib.sdr_pipe_push(self.h, dptr, n, L.SDR_DEVICE_HELD if held else L.SDR_DEVICE))

    def sync(self):
        L.check(L.lib.sdr_pipe_sync(self.h))

    def set_persistent(self, max_session_samples):
        """consume held device vectors with a resident kernel instead of launches (sdr_pipe_set_persistent)"""
        L.check(L.lib.sdr_pipe_set_persistent(self.h, max_session_samples))

    def ready(self):
        n = C.c_int()
        L.check(L.lib.sdr_pipe_ready(self.h, C.byref(n)))
        return n.value

    def pop(self, capacity=None):
        n = C.c_longlong()
        L.check(L.lib.sdr_pipe_next_len(self.h, C.byref(n)))
        out = np.empty(n.value, self.out_dtype)
        L.check(L.lib.sdr_pipe_pop(self.h, L.ptr(out), C.byref(n), L.SDR_HOST))
        return out[:n.value]

    def state_save(self) -> bytes:
        """sdr_pipe_state_save: the stage's carried stream state (tail, counters, carried samples, un-popped outputs)"""
        n = C.c_size_t()
        L.check(L.lib.sdr_pipe_state_size(self.h, C.byref(n)))
        buf = (C.c_char * n.value)()
        w = C.c_size_t()
        L.check(L.lib.sdr_pipe_state_save(self.h, buf, n.value, C.byref(w)))
        return bytes(buf[:w.value])

    def state_restore(self, blob: bytes):
        """sdr_pipe_state_restore into a stage constructed the same way"""
        L.check(L.lib.sdr_pipe_state_restore(self.h, blob, len(blob)))

    def connect(self, dst):
        L.check(L.lib.sdr_pipe_connect(self.h, dst.h))
        self._down = dst   # keep the downstream stage alive as long as this one can forward into it
        return dst

    def close(self):
        if self.h:
            L.lib.sdr_pipe_destroy(self.h)   # also unlinks it from its neighbours (sdr_pipe_destroy)
            self.h = None
        self._down = None

    def __del__(self):
        if L is not None:
            self.close()


def _fir_pipe(create, rec, block_size_out):
    h = C.c_void_p()
    L.check(create(rec.handle, block_size_out, C.byref(h)))
    dt = _dtype(rec.cplx)
    return NativePipe(h, rec.ctx, dt, dt, owner=rec)


def pipeFirFilter(f: Filter, blockSizeOut: int) -> NativePipe:
    return _fir_pipe(L.lib.sdr_pipe_fir_filter, f, blockSizeOut)


def pipeFirDecimator(d: Decimator, blockSizeOut: int) -> NativePipe:
    return _fir_pipe(L.lib.sdr_pipe_fir_decimator, d, blockSizeOut)


def pipeFirResampler(r: Resampler, blockSizeOut: int) -> NativePipe:
    return _fir_pipe(L.lib.sdr_pipe_fir_resampler, r, blockSizeOut)


def _drive(pipe: NativePipe, block_size_out: int, src: Iterable[np.ndarray]) -> Iterator[np.ndarray]:
    try:
        for vec in src:
            pipe.push(vec)
            while pipe.ready():
                yield pipe.pop(block_size_out)
    finally:
        pipe.close()


def firFilter(f: Filter, blockSizeOut: int, src: Iterable[np.ndarray]) -> Iterator[np.ndarray]:
    """firFilter :: Filter -> Int -> Pipe (v a) (v a) m ()   (Filter.hs:532-535)"""
    return _drive(pipeFirFilter(f, blockSizeOut), blockSizeOut, src)


def firDecimator(d: Decimator, blockSizeOut: int, src: Iterable[np.ndarray]) -> Iterator[np.ndarray]:
    """firDecimator (Filter.hs:574-577)"""
    return _drive(pipeFirDecimator(d, blockSizeOut), blockSizeOut, src)


def firResampler(r: Resampler, blockSizeOut: int, src: Iterable[np.ndarray]) -> Iterator[np.ndarray]:
    """firResampler (Filter.hs:679-682)"""
    return _drive(pipeFirResampler(r, blockSizeOut), blockSizeOut, src)
