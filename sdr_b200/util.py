"""Host-side mirror of the convert / scale part of ``SDR.Util`` and of ``SDR.Demod`` for the CUDA backend."""
import ctypes as C
from typing import Iterable, Iterator

import numpy as np

from . import _lib as L
from .filter import NativePipe, default_context


def interleavedIQUnsignedByteToFloat(v) -> np.ndarray:
    """interleavedIQUnsignedByteToFloatFast (Util.hs:137-138) -> convertCAVX (convert.c:37-50): bit-exact"""
    b = np.ascontiguousarray(v, np.uint8)
    if len(b) % 2:
        raise ValueError("interleaved I/Q bytes expected (even length)")
    out = np.empty(len(b), np.float32)
    L.check(L.lib.convertCuda(len(b), L.ptr(b), L.ptr(out)))
    return out.view(np.complex64)


def interleavedIQSigned2048ToFloat(v) -> np.ndarray:
    """interleavedIQSigned2048ToFloatFast (Util.hs:177-178) -> convertCAVXBladeRF (convert.c:73-85): bit-exact"""
    b = np.ascontiguousarray(v, np.int16)
    out = np.empty(len(b), np.float32)
    L.check(L.lib.convertCudaBladeRF(len(b), L.ptr(b), L.ptr(out)))
    return out.view(np.complex64)


def complexFloatToInterleavedIQSigned2048(v) -> np.ndarray:
    """complexFloatToInterleavedIQSigned2048 (Util.hs:202-211) -> convertBladeRFTransmit (convert.c:87-101)"""
    x = L.as_floats(np.asarray(v, np.complex64))
    out = np.empty(len(x), np.int16)
    L.check(L.lib.convertCudaBladeRFTransmit(len(x), L.ptr(x), L.ptr(out)))
    return out


def scaleFast(factor, v) -> np.ndarray:
    """scaleFast (Util.hs:254-255) -> scaleAVX (scale.c:30-36)"""
    x = L.f32(v)
    out = np.empty(len(x), np.float32)
    L.check(L.lib.scaleCuda(len(x), float(factor), L.ptr(x), L.ptr(out)))
    return out


def dcBlocker(v, lastSample=0.0, lastOutput=0.0):
    """dcBlocker (filter.c:152-161) -> (out, finalSample, finalOutput)"""
    x = L.f32(v)
    out = np.empty(len(x), np.float32)
    fs, fo = C.c_float(), C.c_float()
    L.check(L.lib.dcBlockerCuda(len(x), lastSample, lastOutput, C.byref(fs), C.byref(fo), L.ptr(x), L.ptr(out)))
    return out, fs.value, fo.value


def fmDemodVec(last, v) -> np.ndarray:
    """fmDemodVec (Demod.hs:32-36): `last` is the previous buffer's final sample"""
    x = L.as_floats(np.asarray(v, np.complex64))
    n = len(x) // 2
    out = np.empty(n, np.float32)
    last = complex(last)
    L.check(L.lib.fmDemodCuda(n, last.real, last.imag, L.ptr(x), L.ptr(out)))
    return out


def pipeFmDemod(ctx=None) -> NativePipe:
    ctx = ctx or default_context()
    h = C.c_void_p()
    L.check(L.lib.sdr_pipe_fm_demod(ctx.h, C.byref(h)))
    return NativePipe(h, ctx, np.complex64, np.float32)


def pipeConvertU8(ctx=None) -> NativePipe:
    ctx = ctx or default_context()
    h = C.c_void_p()
    L.check(L.lib.sdr_pipe_convert_u8(ctx.h, C.byref(h)))
    return NativePipe(h, ctx, np.uint8, np.complex64)


def pipeScale(factor, ctx=None) -> NativePipe:
    ctx = ctx or default_context()
    h = C.c_void_p()
    L.check(L.lib.sdr_pipe_scale(ctx.h, float(factor), C.byref(h)))
    return NativePipe(h, ctx, np.float32, np.float32)


def pipeDcBlocker(ctx=None) -> NativePipe:
    ctx = ctx or default_context()
    h = C.c_void_p()
    L.check(L.lib.sdr_pipe_dc_blocker(ctx.h, C.byref(h)))
    return NativePipe(h, ctx, np.float32, np.float32)


def dcBlockingFilter(src: Iterable[np.ndarray], ctx=None) -> Iterator[np.ndarray]:
    """dcBlockingFilter :: Pipe (VS.Vector Float) (VS.Vector Float) IO ()  (Filter.hs:730-739)"""
    pipe = pipeDcBlocker(ctx)
    try:
        for vec in src:
            pipe.push(vec)
            yield pipe.pop()
    finally:
        pipe.close()


def pipeFmFrontEnd(decimator, blockSizeOut) -> NativePipe:
    """P.map interleavedIQUnsignedByteToFloat >-> firDecimator decimator blockSizeOut >-> fmDemod  (fm.hs:34-37), fused:
    u8 IQ bytes in, float phases out"""
    h = C.c_void_p()
    L.check(L.lib.sdr_pipe_fm_frontend(decimator.handle, blockSizeOut, C.byref(h)))
    return NativePipe(h, decimator.ctx, np.uint8, np.float32, owner=decimator)


def pipeFmLowRate(resampler, blockSizeResampler, filt, blockSizeOut, scale) -> NativePipe:
    """firResampler resampler blockSizeResampler >-> firFilter filt blockSizeOut >-> P.map (VG.map (* scale))  (fm.hs:38-40),
    fused: float phases in, audio samples out"""
    h = C.c_void_p()
    L.check(L.lib.sdr_pipe_fm_lowrate(resampler.handle, blockSizeResampler, filt.handle, blockSizeOut, float(scale), C.byref(h)))
    return NativePipe(h, resampler.ctx, np.float32, np.float32, owner=(resampler, filt))


def pipeU8Decimator(decimator, blockSizeOut) -> NativePipe:
    """P.map interleavedIQUnsignedByteToFloat >-> firDecimator decimator blockSizeOut  (fm.hs:34-36), fused: u8 IQ bytes
    in, complex samples out"""
    h = C.c_void_p()
    L.check(L.lib.sdr_pipe_u8_decimator(decimator.handle, blockSizeOut, C.byref(h)))
    return NativePipe(h, decimator.ctx, np.uint8, np.complex64, owner=decimator)


def fmDemod(src: Iterable[np.ndarray], ctx=None) -> Iterator[np.ndarray]:
    """fmDemod :: Pipe (v (Complex a)) (v a) IO ()  (Demod.hs:38-46)"""
    pipe = pipeFmDemod(ctx)
    try:
        for vec in src:
            pipe.push(vec)
            yield pipe.pop(len(vec))
    finally:
        pipe.close()
