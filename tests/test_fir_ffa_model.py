"""2-parallel fast-FIR lane function (sdr_b200/csrc/fir_ffa.cuh) checked on the CPU: the product's own
__host__ __device__ function, compiled for the host (tests/emul/ffa_emul.cpp), against the reference's AVX filter
(c_sources/filter.c:37-68 via the pinned oracle port).  The sub-filter sums are associated differently from the
reference's, so the bar is the path's tolerance -- 1e-5 of max(|y_ref|, rms(y_ref)) -- not bit equality; the margin is
asserted too (3e-6)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle
import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "emul", "ffa_emul.cpp")
HDR = os.path.join(ROOT, "sdr_b200", "csrc", "fir_ffa.cuh")
OUT = os.path.join(ROOT, "build", "libffa_emul.so")


@pytest.fixture(scope="module")
def emul():
    inc = "/usr/local/cuda/include"
    if not os.path.exists(os.path.join(inc, "cuda_runtime.h")):
        pytest.skip("cuda_runtime.h not found")
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.run(["g++", "-O2", "-mfma", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", "-I", inc, SRC, "-o", OUT],
                       check=True)
    lib = C.CDLL(OUT)
    lib.emul_fir_ffa.restype = C.c_longlong
    lib.emul_fir_ffa.argtypes = [C.c_int, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p]
    return lib


def ffa(lib, taps, x):
    y = np.full(len(x), np.nan, np.float32)
    done = lib.emul_fir_ffa(len(taps), x.ctypes.data, len(x), taps.ctypes.data, y.ctypes.data)
    assert done > 0
    return y[:done]


@pytest.mark.parametrize("T,cutoff", [(64, 1 / 4), (64, 1 / 20), (32, 1 / 4)])
def test_ffa_matches_reference_avx_filter(emul, port, T, cutoff):
    taps = synth.windowed_sinc_taps(T, cutoff)
    x = synth.noise(200_000)
    got = ffa(emul, taps, x)
    want = port.filter(oracle.V_AVX, len(got), taps, x, False)
    scale = np.maximum(np.abs(want), np.sqrt(np.mean(want.astype(np.float64) ** 2)))
    err = float((np.abs(got - want) / scale).max())
    assert err <= 3e-6, err


def test_ffa_random_taps_and_large_dynamic_range(emul, port):
    rng = np.random.default_rng(3)
    taps = rng.standard_normal(64).astype(np.float32)
    x = (rng.standard_normal(100_000) * np.exp(rng.uniform(-3, 3, 100_000))).astype(np.float32)
    got = ffa(emul, taps, x)
    want = port.filter(oracle.V_AVX, len(got), taps, x, False)
    scale = np.maximum(np.abs(want), np.sqrt(np.mean(want.astype(np.float64) ** 2)))
    assert float((np.abs(got - want) / scale).max()) <= 5e-6
    # exact on integers small enough that no rounding happens anywhere: the decomposition itself is an identity
    ti = rng.integers(-8, 9, 64).astype(np.float32)
    xi = rng.integers(-64, 65, 50_000).astype(np.float32)
    gi = ffa(emul, ti, xi)
    wi = port.filter(oracle.V_AVX, len(gi), ti, xi, False)
    assert np.array_equal(gi, wi)


def test_property_random_taps_and_streams(emul, port):
    """hypothesis, shapes like the reference's QuickCheck generators (values in (-10, 10), symmetric taps): inside the
    path's tolerance against the reference AVX filter, and an identity on integer data"""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=40, deadline=None)
    @given(st.sampled_from([32, 64]), st.integers(200, 20_000), st.integers(0, 2 ** 31 - 1), st.booleans())
    def prop(T, n, seed, integers):
        rng = np.random.default_rng(seed)
        if integers:
            half = rng.integers(-8, 9, T // 2).astype(np.float32)
            x = rng.integers(-64, 65, n).astype(np.float32)
        else:
            half = rng.uniform(-10, 10, T // 2).astype(np.float32)
            x = rng.uniform(-10, 10, n).astype(np.float32)
        taps = np.concatenate([half, half[::-1]])
        got = ffa(emul, taps, x)
        want = port.filter(oracle.V_AVX, len(got), taps, x, False)
        if integers:
            assert np.array_equal(got, want)
        else:
            scale = np.maximum(np.abs(want), np.sqrt(np.mean(want.astype(np.float64) ** 2)))
            assert float((np.abs(got - want) / scale).max()) <= 1e-5

    prop()
