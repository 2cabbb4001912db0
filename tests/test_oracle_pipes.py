"""The oracle's restatement of the reference's Pipes state machines (oracle/pipes.py <- Filter.hs:532-727) against
the closed-form flat-stream model, over ragged input buffers, and the reference's own cross-implementation property
(tests/TestSuite.hs:60-227: every variant agrees within 0.01 absolute)."""
import numpy as np
import pytest

import oracle
from oracle import V_AVX, V_AVX2, V_AVXSYM, V_SCALAR, V_SSE, V_SSE2, V_SSESYM, pipes


def chunks(x, sizes):
    i = 0
    k = 0
    while i < len(x):
        n = sizes[k % len(sizes)]
        yield x[i:i + n]
        i += n
        k += 1


def rel_err(got, want):
    want = np.asarray(want)
    scale = max(np.max(np.abs(want)), 1e-30)
    return np.max(np.abs(got - want)) / scale


def noise(rng, n, cplx):
    if cplx:
        return (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    return rng.standard_normal(n).astype(np.float32)


@pytest.mark.parametrize("sizes", [[8192], [1000, 517, 2048, 300], [256, 4096]])
@pytest.mark.parametrize("block_out", [8192, 1000, 77])
def test_fir_filter_pipe_matches_flat_stream(sizes, block_out):
    rng = np.random.default_rng(1)
    half = rng.standard_normal(32).astype(np.float32)
    taps = np.concatenate([half, half[::-1]])
    x = noise(rng, 40000, False)
    f = pipes.mk_filter_sym_r(V_AVXSYM, half)
    got = list(pipes.fir_filter(f, block_out, chunks(x, sizes)))
    want = pipes.flat_decimate(x, taps, 1)
    y = np.concatenate(got)
    assert all(len(b) == block_out for b in got)
    assert len(want) - len(y) < block_out + 64   # all but the unfinished block and the awaited cross outputs
    assert rel_err(y, want[:len(y)]) < 2e-6


@pytest.mark.parametrize("sizes", [[8192], [1000, 517, 2048, 300], [128, 4096]])
@pytest.mark.parametrize("variant,cplx", [(V_AVX, True), (V_SSE, True), (V_AVX, False), (V_SCALAR, False)])
def test_fir_decimator_pipe_matches_flat_stream(sizes, variant, cplx):
    rng = np.random.default_rng(2)
    taps = (rng.standard_normal(128) / 128).astype(np.float32)
    x = noise(rng, 8192 * 9, cplx)
    d = pipes.mk_decimator_c(variant, 8, taps) if cplx else pipes.mk_decimator(variant, 8, taps)
    got = list(pipes.fir_decimator(d, 1024, chunks(x, sizes)))
    want = pipes.flat_decimate(x, taps, 8)
    y = np.concatenate(got)
    assert all(len(b) == 1024 for b in got)
    assert len(want) - len(y) < 1024 + 16        # everything but the unfinished block (and the awaited tail)
    assert rel_err(y, want[:len(y)]) < 2e-6


def test_fir_decimator_block_counts_cfg2():
    """8192 in -> 1009 (C) + 15 (cross) = 1024 out; one 8192-block yielded per 8 input blocks."""
    rng = np.random.default_rng(3)
    taps = (rng.standard_normal(128) / 128).astype(np.float32)
    calls = []
    d = pipes.mk_decimator_c(V_AVX, 8, taps)
    one, cross = d.decimateOne, d.decimateCross
    d.decimateOne = lambda c, b: (calls.append(("one", c)), one(c, b))[1]
    d.decimateCross = lambda c, l, n: (calls.append(("cross", c)), cross(c, l, n))[1]
    x = noise(rng, 8192 * 17, True)
    got = list(pipes.fir_decimator(d, 8192, chunks(x, [8192])))
    assert len(got) == 2
    assert calls[0] == ("one", 1009) and calls[1] == ("cross", 15) and calls[2] == ("one", 1009)


@pytest.mark.parametrize("L,M,T", [(3, 10, 90), (3, 7, 77), (5, 11, 64), (2, 3, 33)])
@pytest.mark.parametrize("variant,cplx", [(V_AVX, False), (V_SCALAR, False), (V_AVX2, True)])
@pytest.mark.parametrize("sizes", [[8192], [1000, 517, 2048, 300]])
def test_fir_resampler_pipe_matches_flat_stream(L, M, T, variant, cplx, sizes):
    rng = np.random.default_rng(4)
    taps = rng.standard_normal(T).astype(np.float32)
    x = noise(rng, 30000, cplx)
    r = pipes.mk_resampler(variant, L, M, taps, cplx=cplx)
    got = list(pipes.fir_resampler(r, 512, chunks(x, sizes)))
    y = np.concatenate(got)
    want = pipes.flat_resample(x, taps, L, M)
    assert len(y) > 0.9 * len(want) - 512
    assert rel_err(y, want[:len(y)]) < 3e-6


def test_resampler_phase_table_cfg3():
    """SURVEY.md K3 [probe]: 3/10, 90 taps => increments [4,3,3], 30 taps/phase padded to 32, numCoeffsR 96."""
    nc, inc, groups = pipes.prepare_coeffs(8, 3, 10, np.arange(90, dtype=np.float32))
    assert nc == 30 and inc == [4, 3, 3] and groups.shape == (3, 32)
    assert list(groups[:, 0]) == [0.0, 2.0, 1.0]
    assert pipes.mk_resampler(V_AVX, 3, 10, np.arange(90)).numCoeffsR == 96


def test_fm_demod_pipe():
    rng = np.random.default_rng(5)
    x = noise(rng, 5000, True)
    x[17] = 0
    got = np.concatenate(list(pipes.fm_demod(chunks(x, [1000, 333, 2000]))))
    prev = np.concatenate([[0j], x[:-1]]).astype(np.complex128)
    z = x.astype(np.complex128) * np.conj(prev)
    want = np.where(z == 0, 0.0, np.angle(z))
    d = np.abs(got - want)
    d = np.minimum(d, 2 * np.pi - d)
    assert np.max(d) < 1e-6
    assert got[0] == 0 and got[17] == 0 and got[18] == 0


def test_reference_cross_implementation_property(port):
    """TestSuite.hs:74-83,102-110,128-164 restated over the port: all variants within 0.01 absolute of the first."""
    rng = np.random.default_rng(6)
    for size, half_n, factor in [(1024, 32, 1), (2048, 64, 5), (4096, 128, 13)]:
        half = rng.uniform(-10, 10, half_n).astype(np.float32)
        full = np.concatenate([half, half[::-1]])
        dup = pipes.duplicate(full)
        xr, xc = (rng.uniform(-10, 10, size).astype(np.float32),
                  (rng.uniform(-10, 10, size) + 1j * rng.uniform(-10, 10, size)).astype(np.complex64))
        num = (size - 2 * half_n + 1) // factor
        rs = [port.decimate(v, num, factor, c, xr, False) for v, c in
              [(V_SCALAR, full), (V_SSE, full), (V_AVX, full), (V_SSESYM, half), (V_AVXSYM, half)]]
        assert all(np.max(np.abs(r - rs[0])) < 0.01 for r in rs)
        cs = [port.decimate(v, num, factor, c, xc, True) for v, c in
              [(V_SCALAR, full), (V_SSE, dup), (V_AVX, dup), (V_SSE2, full), (V_AVX2, full), (V_SSESYM, half),
               (V_AVXSYM, half)]]
        assert all(np.max(np.abs(c - cs[0])) < 0.01 for c in cs)
        want = pipes.flat_decimate(xc, full, factor, num)
        assert rel_err(cs[2], want) < 2e-6


def test_cross_kernels_sequential_sum(port):
    rng = np.random.default_rng(7)
    taps = rng.standard_normal(64).astype(np.float32)
    last, nxt = noise(rng, 40, True), noise(rng, 200, True)
    got = port.decimate_cross(4, taps, 10, last, nxt, True)
    want = pipes.flat_decimate(np.concatenate([last, nxt]), taps, 4, 10)
    assert rel_err(got, want) < 2e-6
    # order check: exactly the left-to-right float32 sum
    cat = np.concatenate([last, nxt])
    acc = np.float32(0)
    for k in range(64):
        acc = np.float32(acc + np.float32(cat[k].real * taps[k]))
    assert got[0].real == acc
