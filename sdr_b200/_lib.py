"""ctypes binding of ``sdr_b200/lib/libsdr_b200.so`` (the C ABI declared in ``include/sdr_b200.h``).

There is no CPU fallback anywhere in this package: if the shared library is missing the import fails, and if no
sm_100 device is usable every compute call raises :class:`SdrError` with ``SDR_ENODEVICE``.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SDR_B200_LIB_PATH") or os.path.join(_HERE, "lib", "libsdr_b200.so")

SDR_OK, SDR_EINVAL, SDR_EPRECOND, SDR_ECUDA, SDR_ENODEVICE, SDR_ENOMEM, SDR_ENCCL, SDR_EAGAIN = range(8)
SDR_HOST, SDR_DEVICE, SDR_HOST_PINNED, SDR_DEVICE_HELD = 0, 1, 2, 3
SDR_ARITH_FAST, SDR_ARITH_EXACT = 0, 1
V_SCALAR, V_SSE, V_AVX, V_SSE2, V_AVX2, V_SSESYM, V_AVXSYM = range(7)
COMM_ID_BYTES = 128


class SdrError(RuntimeError):
    """Non-zero status from the native library; ``code`` is the SDR_E* value, the message is sdr_last_error()."""

    def __init__(self, code, msg):
        super().__init__(f"[sdr_b200 status {code}] {msg}")
        self.code = code
        self.msg = msg


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `make -C sdr_b200/csrc` (or __graft_entry__.build()); "
        "sdr_b200 has no CPU fallback")

lib = C.CDLL(LIB_PATH)
lib.sdr_last_error.restype = C.c_char_p
lib.sdr_decimator_last_kernel.restype = C.c_char_p
lib.sdr_filter_last_kernel.restype = C.c_char_p
lib.sdr_resampler_last_kernel.restype = C.c_char_p
lib.sdr_filter_last_kernel.argtypes = [C.c_void_p]
lib.sdr_resampler_last_kernel.argtypes = [C.c_void_p]

c_void_pp = C.POINTER(C.c_void_p)
_f32p = C.POINTER(C.c_float)


class ShardPlan(C.Structure):
    """sdr_shard_t"""
    _fields_ = [("in_begin", C.c_longlong), ("in_count", C.c_longlong), ("halo", C.c_longlong),
                ("out_begin", C.c_longlong), ("out_count", C.c_longlong), ("out_interior", C.c_longlong),
                ("n_samples", C.c_longlong), ("taps", C.c_int), ("factor", C.c_int), ("world", C.c_int),
                ("rank", C.c_int)]


class IoStats(C.Structure):
    """sdr_io_stats_t"""
    _fields_ = [("vectors_in", C.c_longlong), ("elements_in", C.c_longlong), ("vectors_out", C.c_longlong),
                ("elements_out", C.c_longlong), ("read_seconds", C.c_double), ("write_seconds", C.c_double)]


SDR_IO_DATAGRAM_IN, SDR_IO_DATAGRAM_OUT = 1, 2


class ResamplerDat(C.Structure):
    """sdr_resampler_dat_t: the Resampler record's existential state (group, offset), Filter.hs:424"""
    _fields_ = [("group", C.c_int), ("offset", C.c_int)]


def _sig(name, *argtypes, restype=C.c_int):
    f = getattr(lib, name)
    f.argtypes = list(argtypes)
    f.restype = restype
    return f


_P, _I, _LL, _SZ, _F = C.c_void_p, C.c_int, C.c_longlong, C.c_size_t, C.c_float

_sig("sdr_device_count", C.POINTER(_I))
_sig("sdr_has_cuda")
_sig("sdr_b200_abi_version")
_sig("sdr_ctx_create", _I, c_void_pp)
_sig("sdr_ctx_destroy", _P)
_sig("sdr_ctx_sync", _P)
_sig("sdr_ctx_set_arith", _P, _I)
_sig("sdr_ctx_set_fast_fir", _P, _I)
_sig("sdr_ctx_sm_count", _P, C.POINTER(_I))
_sig("sdr_dev_alloc", _P, _SZ, c_void_pp)
_sig("sdr_dev_free", _P, _P)
_sig("sdr_host_alloc_pinned", _SZ, c_void_pp)
_sig("sdr_host_free_pinned", _P)
_sig("sdr_memcpy_h2d", _P, _P, _P, _SZ)
_sig("sdr_memcpy_d2h", _P, _P, _P, _SZ)
_sig("sdr_memcpy_d2d", _P, _P, _P, _SZ)
_sig("sdr_memset_dev", _P, _P, _I, _SZ)
_sig("sdr_event_create", _P, c_void_pp)
_sig("sdr_event_record", _P, _P)
_sig("sdr_event_elapsed_ms", _P, _P, C.POINTER(_F))
_sig("sdr_event_destroy", _P)
_sig("sdr_ctx_launch_count", _P, C.POINTER(_LL))

for _n in ("filterCudaRR", "filterCudaSymmetricRR", "filterCudaRC", "filterCudaRCDup", "filterCudaSymmetricRC"):
    _sig(_n, _I, _I, _P, _P, _P)
for _n in ("decimateCudaRR", "decimateCudaSymmetricRR", "decimateCudaRC", "decimateCudaRCDup",
           "decimateCudaSymmetricRC"):
    _sig(_n, _I, _I, _I, _P, _P, _P)
_sig("sdr_exact_decimate", _I, _I, _I, _I, _I, _P, _P, _P)
for _n in ("resampleCudaRR", "resampleCudaRC"):
    _sig(_n, _I, _I, _I, _I, _P, _P, _P, _P)   # returns the next group (>= 0) or -(status)
_sig("sdr_exact_resample", _I, _I, _I, _I, _I, _I, _P, _P, _I, _P, _P, C.POINTER(_I))
_sig("resampleCudaLegacyRR", _I, _I, _I, _I, _I, _P, _P, _P)
_sig("convertCuda", _I, _P, _P)
_sig("convertCudaBladeRF", _I, _P, _P)
_sig("convertCudaBladeRFTransmit", _I, _P, _P)
_sig("scaleCuda", _I, _F, _P, _P)
_sig("dcBlockerCuda", _I, _F, _F, _f32p, _f32p, _P, _P)
_sig("fmDemodCuda", _I, _F, _F, _P, _P)
_sig("sdr_dev_convert_u8", _P, _P, _P, _LL)
_sig("sdr_dev_scale", _P, _F, _P, _P, _LL)
_sig("sdr_dev_fm_demod", _P, _F, _F, _P, _P, _LL)
_sig("sdr_dev_dc_blocker", _P, _F, _F, _P, _P, _LL, _P)
_sig("sdr_dc_blocker_tuning", _P, _I, _I, _I, _LL)
_sig("sdr_dc_blocker_stats", _P, C.POINTER(_LL), C.POINTER(_I))

_sig("sdr_filter_create", _P, _I, _P, _I, _I, c_void_pp)
_sig("sdr_filter_create_sym", _P, _I, _P, _I, c_void_pp)
_sig("sdr_filter_destroy", _P)
_sig("sdr_filter_num_coeffs", _P)
_sig("sdr_filter_one", _P, _I, _P, _P, _I)
_sig("sdr_filter_cross", _P, _I, _P, _I, _P, _I, _P, _I)
_sig("sdr_decimator_create", _P, _I, _I, _P, _I, _I, c_void_pp)
_sig("sdr_decimator_create_sym", _P, _I, _I, _P, _I, c_void_pp)
_sig("sdr_decimator_destroy", _P)
_sig("sdr_decimator_num_coeffs", _P)
_sig("sdr_decimator_factor", _P)
_sig("sdr_decimate_one", _P, _I, _P, _P, _I)
_sig("sdr_decimate_cross", _P, _I, _P, _I, _P, _I, _P, _I)
_sig("sdr_decimate_stream", _P, _P, _LL, _P, _LL)
_sig("sdr_filter_stream", _P, _P, _LL, _P, _LL)
_sig("sdr_resample_stream", _P, _P, _LL, _P, _LL)
lib.sdr_decimator_last_kernel.argtypes = [_P]
_sig("sdr_resampler_create", _P, _I, _I, _I, _P, _I, _I, c_void_pp)
_sig("sdr_resampler_destroy", _P)
_sig("sdr_resampler_num_coeffs", _P)
_sig("sdr_resampler_interpolation", _P)
_sig("sdr_resampler_decimation", _P)
_sig("sdr_resample_one", _P, C.POINTER(ResamplerDat), _I, _P, _P, _I, C.POINTER(_I))
_sig("sdr_resample_cross", _P, C.POINTER(ResamplerDat), _I, _P, _I, _P, _I, _P, _I, C.POINTER(_I))

_sig("sdr_pipe_fir_filter", _P, _I, c_void_pp)
_sig("sdr_pipe_fir_decimator", _P, _I, c_void_pp)
_sig("sdr_pipe_fir_resampler", _P, _I, c_void_pp)
_sig("sdr_pipe_fm_demod", _P, c_void_pp)
_sig("sdr_pipe_fm_frontend", _P, _I, c_void_pp)
_sig("sdr_pipe_u8_decimator", _P, _I, c_void_pp)
_sig("sdr_pipe_fm_lowrate", _P, _I, _P, _I, _F, c_void_pp)
lib.sdr_pipe_last_kernel.restype = C.c_char_p
lib.sdr_pipe_last_kernel.argtypes = [C.c_void_p]
_sig("sdr_pipe_convert_u8", _P, c_void_pp)
_sig("sdr_pipe_scale", _P, _F, c_void_pp)
_sig("sdr_pipe_dc_blocker", _P, c_void_pp)
_sig("sdr_pipe_destroy", _P)
_sig("sdr_pipe_push", _P, _P, _LL, _I)
_sig("sdr_pipe_ready", _P, C.POINTER(_I))
_sig("sdr_pipe_pop", _P, _P, C.POINTER(_LL), _I)
_sig("sdr_pipe_connect", _P, _P)
_sig("sdr_pipe_state_size", _P, C.POINTER(_SZ))
_sig("sdr_pipe_state_save", _P, _P, _SZ, C.POINTER(_SZ))
_sig("sdr_pipe_state_restore", _P, _P, _SZ)
_sig("sdr_pipe_sync", _P)
_sig("sdr_pipe_set_batch", _P, _LL)
_sig("sdr_pipe_set_persistent", _P, _LL)
_sig("sdr_pipe_next_len", _P, C.POINTER(_LL))
_sig("sdr_pipe_run", _P, _P, _P, _LL, _LL, _I, _P, _LL, _I, C.POINTER(_LL))
_sig("sdr_pipe_run_fd", _P, _P, _I, _LL, _LL, _I, _I, C.POINTER(IoStats))

_sig("sdr_shard_plan", _LL, _I, _I, _I, _I, C.POINTER(ShardPlan))
_sig("sdr_comm_unique_id", _P)
_sig("sdr_comm_create", _P, _P, _I, _I, c_void_pp)
_sig("sdr_comm_destroy", _P)
_sig("sdr_comm_barrier", _P)
_sig("sdr_comm_share_chunks", _P, _P)
_sig("sdr_comm_peer_halo_active", _P, _P)
_sig("sdr_decimate_sharded", _P, _P, C.POINTER(ShardPlan), _P, _P)
_sig("sdr_synth_noise", _P, _P, _LL, _LL, C.c_uint32)
_sig("sdr_synth_bytes", _P, _P, _LL, _LL, C.c_uint32)
_sig("sdr_flush_l2", _P)
_sig("sdr_checksum32", _P, _P, _LL, _LL, C.POINTER(C.c_uint64))


def check_group(ret):
    """resampleCuda*: the reference's return convention (next group), negative = -(SDR_E*)"""
    if ret < 0:
        raise SdrError(-ret, lib.sdr_last_error().decode("utf-8", "replace"))
    return ret


def check(status):
    if status != SDR_OK:
        raise SdrError(status, lib.sdr_last_error().decode("utf-8", "replace"))


def ptr(a):
    """data pointer of a C-contiguous numpy array"""
    return a.ctypes.data_as(C.c_void_p)


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def as_floats(x):
    """complex64 -> interleaved float32 view; real -> float32"""
    x = np.ascontiguousarray(x)
    if np.iscomplexobj(x):
        return np.ascontiguousarray(x, dtype=np.complex64).view(np.float32)
    return np.ascontiguousarray(x, dtype=np.float32)
