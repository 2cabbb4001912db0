"""dcBlocker speculation (sdr_b200/csrc/dc_spec.cuh) checked on the CPU: the product's own __host__ __device__ chunk /
verify / repair functions are compiled for the host (tests/emul/dc_emul.cpp) and must reproduce the reference's serial
dcBlocker (c_sources/filter.c:152-161, via the oracle port pinned to it in test_oracle_golden.py) BIT FOR BIT for every
input and every tuning -- good speculation only makes it fast.  The GPU twin of these cases is
tests/test_gpu_parity.py::test_dc_blocker_parallel_*."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "emul", "dc_emul.cpp")
HDR = os.path.join(ROOT, "sdr_b200", "csrc", "dc_spec.cuh")
OUT = os.path.join(ROOT, "build", "libdc_emul.so")


def cuda_include():
    for p in (os.environ.get("CUDA_HOME"), "/usr/local/cuda"):
        if p and os.path.exists(os.path.join(p, "include", "cuda_runtime.h")):
            return os.path.join(p, "include")
    pytest.skip("cuda_runtime.h not found")


@pytest.fixture(scope="module")
def emul():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.run(["g++", "-O2", "-ffp-contract=off", "-fno-fast-math", "-std=c++17", "-shared", "-fPIC", "-I", cuda_include(),
                        SRC, "-o", OUT, "-lm"], check=True)
    lib = C.CDLL(OUT)
    lib.emul_dc_blocker.restype = C.c_int
    lib.emul_dc_blocker.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    return lib


def aligned(n, offset_floats=0):
    raw = np.zeros(n + 16, np.float32)
    skew = (-(raw.ctypes.data // 4)) % 4
    return raw[skew + offset_floats: skew + offset_floats + n]


def run(lib, x, s0=0.0, o0=0.0, ch=2048, k1=6144, k2=4096, vec=2, reverse=0, offset=0, tile=32):
    xin = aligned(len(x), offset)
    xin[:] = x
    out = aligned(len(x), offset)
    out[:] = np.nan
    fin = np.zeros(2, np.float32)
    stats = np.zeros(4, np.uint64)
    assert lib.emul_dc_blocker(xin.ctypes.data, out.ctypes.data, len(x), s0, o0, ch, k1, k2, vec, reverse, fin.ctypes.data,
                               stats.ctypes.data, tile) == 0
    return out, fin, stats


def check(lib, port, x, s0=0.0, o0=0.0, **kw):
    want, ws, wo = port.dc_blocker(x, s0, o0)
    got, fin, stats = run(lib, x, s0, o0, **kw)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), \
        f"{np.count_nonzero(got.view(np.uint32) != want.view(np.uint32))} of {len(x)} words differ ({kw})"
    assert np.array_equal(fin.view(np.uint32), np.array([ws, wo], np.float32).view(np.uint32))
    assert stats[0] == 1 and stats[1] == -(-len(x) // kw.get("ch", 2048))
    return stats


def test_widen_exhaustive_sample(emul):
    """dc_widen (integer float -> double widening, the conversion the device code avoids) == (double)f for every one of
    the 2^32 bit patterns"""
    emul.emul_dc_widen_mismatches.restype = C.c_ulonglong
    emul.emul_dc_widen_mismatches.argtypes = [C.c_ulonglong, C.c_ulonglong]
    assert emul.emul_dc_widen_mismatches(0, 1) == 0


def noise(n, seed=1, scale=1.0, offset=0.0):
    return (np.random.default_rng(seed).standard_normal(n) * scale + offset).astype(np.float32)


def test_default_tuning_on_noise_needs_no_repair(emul, port):
    x = noise(600_000)
    for ch in (1024, 2368, 16384):
        st = check(emul, port, x, 0.25, -0.5, ch=ch)
        assert st[2] == 0, f"chunk {ch}: {st[2]} chunks repaired with the default warm-up"


def test_ragged_lengths_and_chunk_sizes(emul, port):
    for n in (1, 7, 8, 9, 1023, 1024, 1025, 4099, 70_001):
        for ch in (32, 64, 992, 1024):
            check(emul, port, noise(n, seed=n), 0.1, 0.2, ch=ch, k1=64, k2=512)


def test_64_sample_tiles(emul, port):
    """the walk in 256-byte tiles (SDR_B200_DC_TILE=64 on the device): same bits"""
    x = noise(300_001, seed=8)
    check(emul, port, x, 0.25, -0.5, ch=2048, tile=64)
    check(emul, port, x, 0.25, -0.5, ch=4096, k1=0, k2=64, tile=64)
    for n in (1, 63, 64, 65, 4099):
        check(emul, port, x[:n], 0.1, 0.2, ch=64, k1=64, k2=512, tile=64)


def test_scalar_access_path_and_lane_order(emul, port):
    x = noise(50_003, seed=5)
    check(emul, port, x, ch=1024, k1=512, k2=2048, vec=0, offset=1)
    check(emul, port, x, ch=1024, k1=512, k2=2048, vec=0, offset=3, reverse=1)
    check(emul, port, x, ch=1024, k1=512, k2=2048, vec=1, reverse=1)
    check(emul, port, x, ch=1024, k1=512, k2=2048, vec=3)
    check(emul, port, x, ch=1024, k1=512, k2=2048, vec=2)     # `vec` = arithmetic flavour of the exact step (identical bits)


def test_short_warmup_is_repaired(emul, port):
    """with almost no warm-up nearly every chunk misses; the serial repair restores bit-exactness and stops as soon as
    the repaired trajectory meets the stored one"""
    x = noise(200_000, seed=3, scale=3.0, offset=1.0)
    st = check(emul, port, x, 0.0, 7.0, ch=4096, k1=0, k2=32)
    assert st[2] >= 40, st                       # 48 chunks after the first
    assert st[3] < 200_000 - 4096, st            # ... and the repairs merged early at least somewhere
    st = check(emul, port, x, 0.0, 7.0, ch=512, k1=0, k2=0)     # repairs longer than a chunk: carried into the successor
    assert st[2] >= 380, st
    st = check(emul, port, x, ch=4096, k1=6144, k2=512)         # a warm-up that is merely too short: some chunks miss
    assert 0 < st[2] < 48, st


def test_constant_input_denormal_fixed_point(emul, port):
    """exactly constant input parks the true trajectory on a denormal fixed point that speculation from zero never
    reaches: every chunk is repaired, the result is still exact"""
    x = np.full(120_000, 0.5, np.float32)
    st = check(emul, port, x, 0.0, 0.0, ch=2048, k1=1024, k2=1024)
    want, _, wo = port.dc_blocker(x, 0.0, 0.0)
    assert 0 < abs(float(wo)) < 1e-42                      # stuck on a denormal
    assert st[2] > 0
    check(emul, port, np.zeros(100_000, np.float32), ch=2048)            # true trajectory is zero: nothing to repair
    check(emul, port, np.zeros(100_000, np.float32), 0.0, 1.0, ch=2048)  # decays from 1 to the fixed point


def test_special_values(emul, port):
    x = noise(80_000, seed=9)
    x[30_000] = np.inf
    want, _, _ = port.dc_blocker(x)
    got, _, _ = run(emul, x, ch=1024, k1=512, k2=512)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))     # NaN from there on, same payload
    x = noise(80_000, seed=10, scale=1e30)
    check(emul, port, x, ch=1024)
    x = noise(80_000, seed=11, scale=1e-40)                              # denormal inputs
    check(emul, port, x, ch=1024)
    x = (np.random.default_rng(4).integers(0, 256, 90_000).astype(np.float32) - 128) / 128   # converted u8 samples
    st = check(emul, port, x, ch=1024)
    assert st[2] == 0


def test_property_any_input_any_tuning_is_bit_exact(emul, port):
    """hypothesis: whatever the length, the state, the chunk / warm-up lengths, the tile, the arithmetic flavour and the
    input (noise, constant stretches, zeros, denormals, huge values mixed), the chunk-parallel evaluation equals the
    reference recurrence bit for bit"""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.integers(1, 40_000), st.sampled_from([64, 128, 320, 1024, 4096]), st.sampled_from([0, 64, 512, 2048]),
           st.sampled_from([0, 64, 256, 1024]), st.sampled_from([32, 64]), st.integers(0, 3), st.integers(0, 2 ** 31 - 1),
           st.floats(-4, 4, width=32), st.floats(-4, 4, width=32))
    def prop(n, ch, k1, k2, tile, flavour, seed, s0, o0):
        rng = np.random.default_rng(seed)
        x = rng.standard_normal(n).astype(np.float32)
        kind = seed % 5
        if kind == 1:                                   # constant stretches (fixed points of the recurrence)
            x[n // 3: 2 * n // 3] = x[n // 3]
        elif kind == 2:                                 # zeros and denormals
            x[::3] = 0.0
            x[1::7] *= np.float32(1e-42)
        elif kind == 3:                                 # huge dynamic range
            x *= np.exp(rng.uniform(-40, 40, n)).astype(np.float32)
        elif kind == 4:                                 # the u8 sample grid
            x = ((rng.integers(0, 256, n).astype(np.float32) - 128) / 128).astype(np.float32)
        check(emul, port, x, s0, o0, ch=ch, k1=k1, k2=k2, vec=flavour, tile=tile)

    prop()
