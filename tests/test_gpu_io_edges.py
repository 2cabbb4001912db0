"""GPU tests of the producer / consumer edges (sdr_pipe_run_fd): SDR.Serialize.fromHandle / toHandle
(hs_sources/SDR/Serialize.hs:78-83) and SDR.NetworkStream.udpSource / udpSink (hs_sources/SDR/NetworkStream.hs:28-42)
feeding the pipe chain from files, OS pipes and UDP sockets.  What arrives at the other end must be, bit for bit, what
the same chain yields when the vectors are pushed by hand."""
import os
import socket
import struct
import threading

import numpy as np
import pytest

import synth

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(180)]


@pytest.fixture(scope="module")
def sdr():
    import sdr_b200
    assert sdr_b200.has_cuda(), "no sm_100 device: the product has no CPU path"
    return sdr_b200


def by_hand(pipe_factory, vectors):
    head, sink = pipe_factory()
    out = []
    for v in vectors:
        head.push(v)
        while sink.ready():
            out.append(sink.pop())
    return out


def test_file_replay_of_recorded_iq_through_fm_front_end(sdr, tmp_path):
    """fromHandle 16384 h >-> convert >-> firDecimator >-> fmDemod >-> toHandle h' with a recording whose size is not a
    multiple of the vector: the short last vector is pushed too (PB.hGet yields it)"""
    raw = synth.rand_bytes(16384 * 37 + 5000)
    src, dst = tmp_path / "iq.u8", tmp_path / "phase.f32"
    raw.tofile(src)
    taps = synth.windowed_sinc_taps(128, 1 / 16)
    dec = sdr.cudaDecimatorC(8, taps, sizeMultiple=4)

    def chain():
        p = sdr.pipeFmFrontEnd(dec, 1024)
        return p, p
    vecs = [raw[i:i + 16384] for i in range(0, len(raw), 16384)]
    want = by_hand(chain, vecs)
    head, sink = chain()
    with open(src, "rb") as fi, open(dst, "wb") as fo:
        st = sdr.serialize.runHandles(head, sink, 16384, fi, fo)
    got = np.fromfile(dst, np.float32)
    assert st.vectors_in == 38 and st.elements_in == len(raw)
    assert st.vectors_out == len(want) and st.elements_out == 1024 * len(want)
    assert np.array_equal(got.view(np.uint32), np.concatenate(want).view(np.uint32))
    # maxVectors stops early; output discarded
    head, sink = chain()
    with open(src, "rb") as fi:
        st = sdr.serialize.runHandles(head, sink, 16384, fi, None, maxVectors=10)
    assert st.vectors_in == 10 and st.elements_out == (10 * 8192 - 128) // 8 // 1024 * 1024


def test_short_last_vector_raises_the_reference_assert_after_flushing(sdr, tmp_path):
    """a last vector shorter than numCoeffs trips `decimate 1` (Filter.hs:586) -- after everything before it was written"""
    x = synth.noise_complex(8192 * 3 + 50)
    src, dst = tmp_path / "iq.c64", tmp_path / "out.c64"
    x.tofile(src)
    dec = sdr.cudaDecimatorC(8, synth.windowed_sinc_taps(128, 1 / 16), sizeMultiple=4)
    want = np.concatenate(by_hand(lambda: (lambda p: (p, p))(sdr.pipeFirDecimator(dec, 512)), [x[i * 8192:(i + 1) * 8192] for i in range(3)]))
    p = sdr.pipeFirDecimator(dec, 512)
    with open(src, "rb") as fi, open(dst, "wb") as fo:
        with pytest.raises(sdr.SdrError) as e:
            sdr.serialize.runHandles(p, p, 8192, fi, fo)
    assert e.value.code == sdr._lib.SDR_EPRECOND and "decimate 1" in e.value.msg
    got = np.fromfile(dst, np.complex64)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_os_pipe_source_with_partial_reads(sdr):
    """a producer that writes in odd-sized pieces: fromHandle still delivers whole vectors (hGet semantics)"""
    x = synth.noise(4096 * 21)
    half = synth.windowed_sinc_taps(64, 1 / 4)[:32]
    fil = sdr.cudaFilterSymR(half)

    def chain():
        a = sdr.pipeFirFilter(fil, 4096)
        b = sdr.pipeScale(0.2, a.ctx)
        a.connect(b)
        return a, b
    want = np.concatenate(by_hand(chain, [x[i:i + 4096] for i in range(0, len(x), 4096)]))
    r, w = os.pipe()
    r2, w2 = os.pipe()

    def produce():
        b, i, k = x.tobytes(), 0, 0
        try:
            while i < len(b):
                step = (1000, 7, 16384, 333, 65536)[k % 5]
                os.write(w, b[i:i + step])
                i += step
                k += 1
        except OSError:
            pass   # the reader went away (a failed run): the assertion below reports it
        finally:
            os.close(w)
    got = []

    def consume():
        while True:
            b = os.read(r2, 1 << 20)
            if not b:
                break
            got.append(b)
    tp, tc = threading.Thread(target=produce), threading.Thread(target=consume)
    tp.start(); tc.start()
    head, sink = chain()
    try:
        st = sdr.serialize.runHandles(head, sink, 4096, r, w2)
    finally:
        # close our ends first: if the run failed early the producer must get EPIPE and the consumer EOF, never block
        os.close(r); os.close(w2)
        tp.join(30); tc.join(30)
        os.close(r2)
    out = np.frombuffer(b"".join(got), np.float32)
    assert st.vectors_in == 21
    assert np.array_equal(out.view(np.uint32), want.view(np.uint32))


def test_udp_source_and_sink(sdr):
    """udpSource sock size >-> fmDemod >-> udpSink: one datagram = one vector in, one vector = one datagram out"""
    rx = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
    rx.bind(("127.0.0.1", 0))
    rx.setsockopt(socket.SOL_SOCKET, socket.SO_RCVBUF, 1 << 22)
    rx.setsockopt(socket.SOL_SOCKET, socket.SO_RCVTIMEO, struct.pack("ll", 10, 0))   # a lost datagram fails the read, never hangs
    back = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
    back.bind(("127.0.0.1", 0))
    back.setsockopt(socket.SOL_SOCKET, socket.SO_RCVBUF, 1 << 22)
    back.settimeout(10)
    out = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
    out.connect(back.getsockname())
    tx = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
    x = synth.noise_complex(1024 * 12)
    lens = [1024, 512, 1024, 100, 1024, 1024, 1, 1024, 1024, 1024, 1024, 1024]
    vecs, i = [], 0
    for n in lens:
        vecs.append(x[i:i + n]); i += n
    for v in vecs:                      # the datagrams wait in the socket buffer
        tx.sendto(v.tobytes(), rx.getsockname())
    want = by_hand(lambda: (lambda p: (p, p))(sdr.pipeFmDemod()), vecs)
    p = sdr.pipeFmDemod()
    st = sdr.serialize.runUdp(p, p, rx, 1024 * 8, len(vecs), out)
    assert st.vectors_in == len(vecs) and st.vectors_out == len(vecs)
    for w in want:
        d = np.frombuffer(back.recv(65536), np.float32)
        assert np.array_equal(d.view(np.uint32), w.view(np.uint32))
    for s in (rx, back, out, tx):
        s.close()


@pytest.mark.parametrize("flag", [[], ["--reference-coeffs"]])
def test_fm_receiver_example_end_to_end(tmp_path, flag):
    """examples/fm_receiver.py (the reference's examples/fm/fm.hs on the device, two fused stages, file in / file out) ==
    the six stages of fm.hs:34-40 pushed one after the other from host vectors, bit for bit; with the BASELINE shapes and
    with the reference example's own coefficient sets"""
    import subprocess
    import sys
    import sdr_b200
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    rng = np.random.default_rng(21)
    n_vec = 700
    raw = rng.integers(0, 256, 16384 * n_vec, dtype=np.uint8)
    fin, fout = tmp_path / "iq.u8", tmp_path / "audio.f32"
    raw.tofile(fin)
    r = subprocess.run([sys.executable, os.path.join(root, "examples", "fm_receiver.py"), str(fin), str(fout)] + flag,
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    got = np.fromfile(fout, np.float32)
    if flag:
        fm = np.load(os.path.join(root, "tests", "golden", "fm_example_coeffs.npz"))
        c_rf, c_rs, c_au = fm["coeffsRFDecim"], fm["coeffsAudioResampler"], fm["coeffsAudioFilter"]
    else:
        w = sdr_b200.windowed_sinc_taps
        c_rf, c_rs, c_au = w(128, 1 / 16), w(90, 1 / 20, gain=3.0), w(64, 1 / 4)[:32]
    q = [sdr_b200.pipeConvertU8(), sdr_b200.pipeFirDecimator(sdr_b200.cudaDecimatorC(8, c_rf, sizeMultiple=4), 8192), sdr_b200.pipeFmDemod(),
         sdr_b200.pipeFirResampler(sdr_b200.cudaResamplerR(3, 10, c_rs, sizeMultiple=8), 8192),
         sdr_b200.pipeFirFilter(sdr_b200.cudaFilterSymR(c_au), 8192), sdr_b200.pipeScale(0.2)]
    for a, b in zip(q, q[1:]):
        a.connect(b)
    want = []
    for i in range(n_vec):
        q[0].push(raw[16384 * i:16384 * (i + 1)])
        while q[-1].ready():
            want.append(q[-1].pop())
    want = np.concatenate(want)
    for p in q:
        p.close()
    assert len(got) == len(want) and len(got) >= 8192 * 20
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
