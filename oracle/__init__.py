"""TEST INFRASTRUCTURE ONLY -- Python access to the CPU checkers.

* ``port``  : ctypes wrappers over ``oracle/liboracle.so`` (oracle/sdr_oracle.c, our plain-C restatement).
* ``ref``   : ctypes wrappers over ``oracle/_ref/libsdrref.so`` -- the UNMODIFIED reference C (adamwalker/sdr
              c_sources/*.c) compiled by oracle/Makefile; ``None`` when the .so is absent.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this package.  Nothing under ``sdr_b200/`` does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

V_SCALAR, V_SSE, V_AVX, V_SSE2, V_AVX2, V_SSESYM, V_AVXSYM = range(7)

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_i16p = np.ctypeslib.ndpointer(dtype=np.int16, flags="C_CONTIGUOUS")


def build():
    """Compile liboracle.so (and _ref/libsdrref.so when /root/reference is present)."""
    subprocess.run(["make", "-C", _HERE, "all"], check=True, stdout=subprocess.DEVNULL)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def as_floats(x):
    """complex64 -> interleaved float32 view; float32 passes through."""
    x = np.ascontiguousarray(x)
    if np.iscomplexobj(x):
        return np.ascontiguousarray(x.astype(np.complex64)).view(np.float32)
    return x.astype(np.float32, copy=False)


class _Port:
    """Wrappers over liboracle.so. Function names follow oracle/sdr_oracle.c."""

    def __init__(self):
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = self.lib = C.CDLL(path)
        L.o_decimateR.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _f32p, _f32p, _f32p]
        L.o_decimateC.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _f32p, _f32p, _f32p]
        L.o_dcBlocker.argtypes = [C.c_int, C.c_float, C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                  _f32p, _f32p]
        L.o_resampleRR.argtypes = [C.c_int] * 5 + [_f32p, _f32p, _f32p]
        for f in (L.o_resampleR, L.o_resampleC):
            f.argtypes = [C.c_int] * 5 + [_i32p, _f32p, C.c_int, _f32p, _f32p]
            f.restype = C.c_int
        for f in (L.o_decimateCrossR, L.o_decimateCrossC):
            f.argtypes = [C.c_int, C.c_int, _f32p, C.c_int, _f32p, C.c_int, _f32p, C.c_int, _f32p]
        for f in (L.o_resampleCrossR, L.o_resampleCrossC):
            f.argtypes = [C.c_int, C.c_int, C.c_int, _f32p, C.c_int, C.c_int, _f32p, C.c_int, _f32p, C.c_int, _f32p]
            f.restype = C.c_int
        L.o_convertC.argtypes = [C.c_int, _u8p, _f32p]
        L.o_convertCBladeRF.argtypes = [C.c_int, _i16p, _f32p]
        L.o_convertBladeRFTransmit.argtypes = [C.c_int, _f32p, _i16p]
        L.o_scale.argtypes = [C.c_int, C.c_float, _f32p, _f32p]
        L.o_fmDemod.argtypes = [C.c_int, C.c_float, C.c_float, _f32p, _f32p]

    # ---- filters / decimators -------------------------------------------------------------------------------
    def decimate(self, variant, num, factor, coeffs, x, complex_data):
        """coeffs exactly as the reference function of that variant receives them (dup / half / plain)."""
        coeffs = _f32(coeffs)
        xin = as_floats(x)
        if complex_data:
            out = np.zeros(2 * num, np.float32)
            self.lib.o_decimateC(variant, num, factor, len(coeffs), coeffs, xin, out)
            return out.view(np.complex64)
        out = np.zeros(num, np.float32)
        self.lib.o_decimateR(variant, num, factor, len(coeffs), coeffs, xin, out)
        return out

    def filter(self, variant, num, coeffs, x, complex_data):
        return self.decimate(variant, num, 1, coeffs, x, complex_data)

    def dc_blocker(self, x, last_sample=0.0, last_output=0.0):
        x = _f32(x)
        out = np.zeros(len(x), np.float32)
        fs, fo = C.c_float(), C.c_float()
        self.lib.o_dcBlocker(len(x), last_sample, last_output, C.byref(fs), C.byref(fo), x, out)
        return out, fs.value, fo.value

    # ---- resamplers -----------------------------------------------------------------------------------------
    def resample_legacy(self, num, interpolation, decimation, offset, coeffs, x):
        coeffs, x = _f32(coeffs), _f32(x)
        out = np.zeros(num, np.float32)
        self.lib.o_resampleRR(num, len(coeffs), interpolation, decimation, offset, coeffs, x, out)
        return out

    def resample_n(self, variant, num, num_coeffs, starting_group, increments, groups, x, complex_data):
        """groups: float32 [num_groups, padded taps]; num_coeffs is what the reference passes to C: the UNPADDED
        maximum group length (FilterInternal.hs:300,340) -- the SIMD loop over-runs into the zero padding."""
        groups = np.ascontiguousarray(groups, np.float32)
        inc = np.ascontiguousarray(increments, np.int32)
        xin = as_floats(x)
        if complex_data:
            out = np.zeros(2 * num, np.float32)
            g = self.lib.o_resampleC(variant, num, num_coeffs, starting_group, groups.shape[0], inc, groups,
                                     groups.shape[1], xin, out)
            return out.view(np.complex64), g
        out = np.zeros(num, np.float32)
        g = self.lib.o_resampleR(variant, num, num_coeffs, starting_group, groups.shape[0], inc, groups,
                                 groups.shape[1], xin, out)
        return out, g

    # ---- cross-buffer kernels ------------------------------------------------------------------------------
    def decimate_cross(self, factor, coeffs, num, last, nxt, complex_data):
        coeffs = _f32(coeffs)
        l, n = as_floats(last), as_floats(nxt)
        div = 2 if complex_data else 1
        nl, nn = len(l) // div, len(n) // div
        if len(l) == 0:
            l = np.zeros(2, np.float32)
        out = np.zeros(num * div, np.float32)
        f = self.lib.o_decimateCrossC if complex_data else self.lib.o_decimateCrossR
        f(factor, len(coeffs), coeffs, num, l, nl, n, nn, out)
        return out.view(np.complex64) if complex_data else out

    def resample_cross(self, interpolation, decimation, coeffs, filter_offset, count, last, nxt, complex_data):
        coeffs = _f32(coeffs)
        l, n = as_floats(last), as_floats(nxt)
        div = 2 if complex_data else 1
        nl, nn = len(l) // div, len(n) // div
        if len(l) == 0:
            l = np.zeros(2, np.float32)
        out = np.zeros(count * div, np.float32)
        f = self.lib.o_resampleCrossC if complex_data else self.lib.o_resampleCrossR
        off = f(interpolation, decimation, len(coeffs), coeffs, filter_offset, count, l, nl, n, nn, out)
        return (out.view(np.complex64) if complex_data else out), off

    # ---- converts / scale / demod ---------------------------------------------------------------------------
    def convert_u8(self, x):
        x = np.ascontiguousarray(x, np.uint8)
        out = np.zeros(len(x), np.float32)
        self.lib.o_convertC(len(x), x, out)
        return out

    def convert_i16(self, x):
        x = np.ascontiguousarray(x, np.int16)
        out = np.zeros(len(x), np.float32)
        self.lib.o_convertCBladeRF(len(x), x, out)
        return out

    def convert_tx(self, x):
        x = _f32(x)
        out = np.zeros(len(x), np.int16)
        self.lib.o_convertBladeRFTransmit(len(x), x, out)
        return out

    def scale(self, factor, x):
        x = _f32(x)
        out = np.zeros(len(x), np.float32)
        self.lib.o_scale(len(x), factor, x, out)
        return out

    def fm_demod(self, x, last=0j):
        xin = as_floats(np.asarray(x, np.complex64))
        n = len(xin) // 2
        out = np.zeros(n, np.float32)
        last = complex(last)
        self.lib.o_fmDemod(n, last.real, last.imag, xin, out)
        return out


class _Ref:
    """Wrappers over the compiled, unmodified reference C. Symbol names are the reference's own."""

    FILTERS_R = {"filterRR": V_SCALAR, "filterSSERR": V_SSE, "filterAVXRR": V_AVX,
                 "filterSSESymmetricRR": V_SSESYM, "filterAVXSymmetricRR": V_AVXSYM}
    FILTERS_C = {"filterRC": V_SCALAR, "filterSSERC": V_SSE, "filterAVXRC": V_AVX, "filterSSERC2": V_SSE2,
                 "filterAVXRC2": V_AVX2, "filterSSESymmetricRC": V_SSESYM, "filterAVXSymmetricRC": V_AVXSYM}
    DECIM_R = {"decimateRR": V_SCALAR, "decimateSSERR": V_SSE, "decimateAVXRR": V_AVX,
               "decimateSSESymmetricRR": V_SSESYM, "decimateAVXSymmetricRR": V_AVXSYM}
    DECIM_C = {"decimateRC": V_SCALAR, "decimateSSERC": V_SSE, "decimateAVXRC": V_AVX, "decimateSSERC2": V_SSE2,
               "decimateAVXRC2": V_AVX2, "decimateSSESymmetricRC": V_SSESYM, "decimateAVXSymmetricRC": V_AVXSYM}
    RESAMP_R = {"resample2RR": V_SCALAR, "resampleSSERR": V_SSE, "resampleAVXRR": V_AVX}
    RESAMP_C = {"resample2RC": V_SCALAR, "resampleSSERC": V_SSE2, "resampleAVXRC": V_AVX2}

    def __init__(self, path):
        L = self.lib = C.CDLL(path)
        for n in list(self.FILTERS_R) + list(self.FILTERS_C):
            getattr(L, n).argtypes = [C.c_int, C.c_int, _f32p, _f32p, _f32p]
            getattr(L, n).restype = None
        for n in list(self.DECIM_R) + list(self.DECIM_C):
            getattr(L, n).argtypes = [C.c_int, C.c_int, C.c_int, _f32p, _f32p, _f32p]
            getattr(L, n).restype = None
        for n in list(self.RESAMP_R) + list(self.RESAMP_C):
            getattr(L, n).argtypes = [C.c_int] * 4 + [_i32p, C.POINTER(C.POINTER(C.c_float)), _f32p, _f32p]
            getattr(L, n).restype = C.c_int
        L.resampleRR.argtypes = [C.c_int] * 5 + [_f32p, _f32p, _f32p]
        L.dcBlocker.argtypes = [C.c_int, C.c_float, C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float), _f32p,
                                _f32p]
        for n in ("convertC", "convertCSSE", "convertCAVX"):
            getattr(L, n).argtypes = [C.c_int, _u8p, _f32p]
        for n in ("convertCBladeRF", "convertCSSEBladeRF", "convertCAVXBladeRF"):
            getattr(L, n).argtypes = [C.c_int, _i16p, _f32p]
        L.convertBladeRFTransmit.argtypes = [C.c_int, _f32p, _i16p]
        for n in ("scale", "scaleSSE", "scaleAVX"):
            getattr(L, n).argtypes = [C.c_int, C.c_float, _f32p, _f32p]

    def filter(self, name, num, coeffs, x):
        cplx = name in self.FILTERS_C
        coeffs, xin = _f32(coeffs), as_floats(x)
        out = np.zeros(num * (2 if cplx else 1), np.float32)
        getattr(self.lib, name)(num, len(coeffs), coeffs, xin, out)
        return out.view(np.complex64) if cplx else out

    def decimate(self, name, num, factor, coeffs, x):
        cplx = name in self.DECIM_C
        coeffs, xin = _f32(coeffs), as_floats(x)
        out = np.zeros(num * (2 if cplx else 1), np.float32)
        getattr(self.lib, name)(num, factor, len(coeffs), coeffs, xin, out)
        return out.view(np.complex64) if cplx else out

    def resample(self, name, num, num_coeffs, starting_group, increments, groups, x):
        cplx = name in self.RESAMP_C
        groups = np.ascontiguousarray(groups, np.float32)
        rows = (C.POINTER(C.c_float) * groups.shape[0])(
            *[groups[g].ctypes.data_as(C.POINTER(C.c_float)) for g in range(groups.shape[0])])
        inc = np.ascontiguousarray(increments, np.int32)
        xin = as_floats(x)
        out = np.zeros(num * (2 if cplx else 1), np.float32)
        g = getattr(self.lib, name)(num, num_coeffs, starting_group, groups.shape[0], inc, rows, xin, out)
        return (out.view(np.complex64) if cplx else out), g

    def resample_legacy(self, num, interpolation, decimation, offset, coeffs, x):
        coeffs, x = _f32(coeffs), _f32(x)
        out = np.zeros(num, np.float32)
        self.lib.resampleRR(num, len(coeffs), interpolation, decimation, offset, coeffs, x, out)
        return out

    def dc_blocker(self, x, last_sample=0.0, last_output=0.0):
        x = _f32(x)
        out = np.zeros(len(x), np.float32)
        fs, fo = C.c_float(), C.c_float()
        self.lib.dcBlocker(len(x), last_sample, last_output, C.byref(fs), C.byref(fo), x, out)
        return out, fs.value, fo.value

    def convert_u8(self, name, x):
        x = np.ascontiguousarray(x, np.uint8)
        out = np.zeros(len(x), np.float32)
        getattr(self.lib, name)(len(x), x, out)
        return out

    def convert_i16(self, name, x):
        x = np.ascontiguousarray(x, np.int16)
        out = np.zeros(len(x), np.float32)
        getattr(self.lib, name)(len(x), x, out)
        return out

    def convert_tx(self, x):
        x = _f32(x)
        out = np.zeros(len(x), np.int16)
        self.lib.convertBladeRFTransmit(len(x), x, out)
        return out

    def scale(self, name, factor, x):
        x = _f32(x)
        out = np.zeros(len(x), np.float32)
        getattr(self.lib, name)(len(x), factor, x, out)
        return out


_port = None
_ref = False


def port():
    global _port
    if _port is None:
        _port = _Port()
    return _port


def ref():
    """The compiled reference, or None when oracle/_ref/libsdrref.so does not exist."""
    global _ref
    if _ref is False:
        path = os.path.join(_HERE, "_ref", "libsdrref.so")
        _ref = _Ref(path) if os.path.exists(path) else None
    return _ref
