"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol include/sdr_b200.h declares,
the compute entry points fail loudly without a device, and the pure-host shard planner is right."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "sdr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{}]*\)\s*;", src)
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    import sdr_b200
    lib = C.CDLL(sdr_b200.LIB_PATH)
    names = _declared_functions()
    assert len(names) > 80, names
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_no_cpu_fallback_without_device():
    import sdr_b200
    if sdr_b200.has_cuda():
        pytest.skip("a GPU is present")
    with pytest.raises(sdr_b200.SdrError) as e:
        sdr_b200.Context(0)
    assert e.value.code == 4  # SDR_ENODEVICE
    with pytest.raises(sdr_b200.SdrError):
        sdr_b200.interleavedIQUnsignedByteToFloat(np.zeros(16, np.uint8))
    with pytest.raises(sdr_b200.SdrError):
        sdr_b200.scaleFast(2.0, np.zeros(16, np.float32))


def test_feature_select_is_first_match_with_default():
    """featureSelect (CPUID.hs:100-104): first matching (predicate, implementation) pair, else the default; hasCUDA is
    the predicate this backend adds in front of the reference's hasAVX / hasSSE42"""
    import sdr_b200
    info = {"avx": True, "sse42": True}
    opts = [(lambda i: False, "cuda"), (lambda i: i["avx"], "avx"), (lambda i: i["sse42"], "sse")]
    assert sdr_b200.featureSelect(info, "c", opts) == "avx"
    assert sdr_b200.featureSelect({"avx": False, "sse42": False}, "c", opts) == "c"
    assert sdr_b200.featureSelect(info, "c", []) == "c"
    want = "cuda" if sdr_b200.has_cuda() else "avx"
    assert sdr_b200.featureSelect(info, "c", [(sdr_b200.hasCUDA, "cuda")] + opts[1:]) == want


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "sdr_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "liboracle" not in text and "libsdrref" not in text, f


@pytest.mark.parametrize("n,taps,factor,world", [(2 ** 28, 128, 8, 8), (2 ** 28, 128, 8, 4), (2 ** 28, 128, 8, 2),
                                                 (2 ** 28, 128, 8, 1), (100000, 51, 3, 4), (8192 * 5 + 17, 64, 1, 3),
                                                 (5000, 128, 8, 8), (300, 128, 8, 2)])
def test_shard_plan_partitions_outputs(n, taps, factor, world):
    from sdr_b200.multigpu import shard_plan
    total_out = (n - taps) // factor + 1 if n >= taps else 0
    nxt_in, nxt_out = 0, 0
    for r in range(world):
        p = shard_plan(n, taps, factor, world, r)
        assert p.in_begin == nxt_in and p.in_begin % 16 == 0 or p.in_count == 0
        nxt_in = p.in_begin + p.in_count
        assert p.out_begin == nxt_out or p.out_count == 0
        if p.out_count:
            nxt_out = p.out_begin + p.out_count
            # every owned window starts inside the chunk; halo covers exactly the overrun of the last one
            assert p.in_begin <= p.out_begin * factor < p.in_begin + p.in_count
            last_end = (p.out_begin + p.out_count - 1) * factor + taps
            assert p.halo == max(0, last_end - (p.in_begin + p.in_count))
            assert p.halo <= taps - 1
            assert 0 <= p.out_interior <= p.out_count
            # interior windows end inside the chunk, the first boundary window does not
            if p.out_interior:
                assert (p.out_begin + p.out_interior - 1) * factor + taps <= p.in_begin + p.in_count
            if p.out_interior < p.out_count:
                assert (p.out_begin + p.out_interior) * factor + taps > p.in_begin + p.in_count
        if r == world - 1:
            assert p.halo == 0
    assert nxt_in == n
    assert nxt_out == total_out


def test_synth_replica_statistics():
    import synth
    x = synth.noise(1 << 16)
    assert abs(float(x.mean())) < 0.02 and abs(float(x.std()) - 1.0) < 0.02
    assert np.array_equal(synth.noise(100, first=50), synth.noise(150)[50:])
    b = synth.rand_bytes(1 << 16)
    assert b.min() == 0 and b.max() == 255
    assert np.array_equal(synth.rand_bytes(64, first=7), synth.rand_bytes(71)[7:])
    t = synth.windowed_sinc_taps(128, 1 / 16)
    assert np.allclose(t, t[::-1]) and abs(float(t.sum()) - 1.0) < 0.01
