"""Stream state export / import (sdr_pipe_state_save / _restore): run half a stream, save, destroy the stage, restore into
a NEW stage built the same way, finish -- the output must be bit-identical to the uninterrupted run, for every stage kind.
The state is the reference's own carried state: the crossover tail (Filter.hs:558-569, 600-611, 712-727), the resampler's
(group, offset) (Filter.hs:419-424), fmDemod's last sample (Demod.hs:41-46), the dcBlocker pair (Filter.hs:731-739)."""
import numpy as np
import pytest

import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sdr():
    import sdr_b200
    assert sdr_b200.has_cuda()
    return sdr_b200


def _noise(n, cplx, seed):
    rng = np.random.default_rng(seed)
    if cplx:
        return (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    return rng.standard_normal(n).astype(np.float32)


def _stages(sdr):
    taps128 = synth.windowed_sinc_taps(128, 1 / 16)
    half = synth.windowed_sinc_taps(64, 1 / 4)[:32]
    t90 = synth.windowed_sinc_taps(90, 1 / 20, gain=3.0)
    rng = np.random.default_rng(5)
    t51 = (rng.standard_normal(51) / 7).astype(np.float32)
    # name -> (factory returning (pipe, owners...), input dtype maker, sizes)
    return {
        "firFilter": (lambda: sdr.pipeFirFilter(sdr.cudaFilterSymR(half), 1000), lambda n: _noise(n, False, 1)),
        "firDecimator": (lambda: sdr.pipeFirDecimator(sdr.cudaDecimatorC(8, taps128, sizeMultiple=4), 512), lambda n: _noise(n, True, 2)),
        "firDecimator51": (lambda: sdr.pipeFirDecimator(sdr.cudaDecimatorC(8, t51, sizeMultiple=4), 300), lambda n: _noise(n, True, 3)),
        "firResampler": (lambda: sdr.pipeFirResampler(sdr.cudaResamplerR(3, 10, t90, sizeMultiple=8), 700), lambda n: _noise(n, False, 4)),
        "firResamplerC": (lambda: sdr.pipeFirResampler(sdr.cudaResamplerC(3, 10, t90, sizeMultiple=4), 700), lambda n: _noise(n, True, 5)),
        "fmDemod": (lambda: sdr.pipeFmDemod(), lambda n: _noise(n, True, 6)),
        "dcBlocker": (lambda: sdr.pipeDcBlocker(), lambda n: _noise(n, False, 7) + np.float32(0.3)),
        "fmFrontEnd": (lambda: sdr.pipeFmFrontEnd(sdr.cudaDecimatorC(8, taps128, sizeMultiple=4), 400),
                       lambda n: np.random.default_rng(8).integers(0, 256, 2 * n, dtype=np.uint8)),
        "u8Decimator": (lambda: sdr.pipeU8Decimator(sdr.cudaDecimatorC(8, taps128, sizeMultiple=4), 400),
                        lambda n: np.random.default_rng(9).integers(0, 256, 2 * n, dtype=np.uint8)),
        "scale": (lambda: sdr.pipeScale(0.2), lambda n: _noise(n, False, 10)),
    }


SIZES = [8192, 3001, 70000, 1300, 8192, 131072 + 6, 2000, 9000]


def _drain(pipe, out):
    while pipe.ready():
        out.append(pipe.pop())


@pytest.mark.parametrize("kind", ["firFilter", "firDecimator", "firDecimator51", "firResampler", "firResamplerC", "fmDemod", "dcBlocker",
                                  "fmFrontEnd", "u8Decimator", "scale"])
@pytest.mark.parametrize("cut", [3, 5])
def test_save_destroy_restore_continues_bit_for_bit(sdr, kind, cut):
    make, data = _stages(sdr)[kind]
    byte_fed = kind in ("fmFrontEnd", "u8Decimator")
    x = data(sum(SIZES))
    k = 2 if byte_fed else 1
    vecs, o = [], 0
    for n in SIZES:
        vecs.append(x[k * o:k * (o + n)])
        o += n
    # uninterrupted
    p = make()
    want = []
    for v in vecs:
        p.push(v)
        _drain(p, want)
    p.close()
    # interrupted after `cut` vectors; the last pushed vector's outputs are deliberately left un-popped in the stage
    p = make()
    got = []
    for i, v in enumerate(vecs[:cut]):
        p.push(v)
        if i + 1 < cut:
            _drain(p, got)
    blob = p.state_save()
    p.close()
    del p
    q = make()
    with pytest.raises(sdr.SdrError):
        q.state_restore(blob[:40])            # truncated
    q.state_restore(blob)
    _drain(q, got)
    for v in vecs[cut:]:
        q.push(v)
        _drain(q, got)
    q.close()
    assert len(got) == len(want) and len(want) > 0
    for a, b in zip(got, want):
        assert a.dtype == b.dtype and a.shape == b.shape
        assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), kind


def test_restore_rejects_a_differently_built_stage(sdr):
    taps = synth.windowed_sinc_taps(128, 1 / 16)
    a = sdr.pipeFirDecimator(sdr.cudaDecimatorC(8, taps, sizeMultiple=4), 512)
    a.push(_noise(8192, True, 1))
    blob = a.state_save()
    for other in (sdr.pipeFirDecimator(sdr.cudaDecimatorC(4, taps, sizeMultiple=4), 512),
                  sdr.pipeFirDecimator(sdr.cudaDecimatorC(8, taps, sizeMultiple=4), 256),
                  sdr.pipeFirFilter(sdr.cudaFilterC(taps), 512)):
        with pytest.raises(sdr.SdrError) as e:
            other.state_restore(blob)
        assert e.value.code == 1   # SDR_EINVAL
        other.close()
    a.close()
