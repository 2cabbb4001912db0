"""Host-side mirror of the producer / consumer edges either side of the path: ``SDR.Serialize.fromHandle`` / ``toHandle``
(hs_sources/SDR/Serialize.hs:78-83) and ``SDR.NetworkStream.udpSource`` / ``udpSink`` (hs_sources/SDR/NetworkStream.hs:28-42),
run natively: ``read()`` lands in a page-locked ring and the vectors go through the pipe chain without touching Python."""
import ctypes as C

from . import _lib as L


def _fd(handle):
    return handle if isinstance(handle, int) else handle.fileno()


def runHandles(head, sink, samples, inHandle, outHandle=None, maxVectors=0):
    """``runEffect $ fromHandle samples inHandle >-> head >-> ... >-> sink >-> toHandle outHandle``: vectors of exactly
    `samples` input elements until end of file (or maxVectors); outHandle None discards the output (devnull,
    PipeUtils.hs:36).  Returns the sdr_io_stats_t counters."""
    st = L.IoStats()
    L.check(L.lib.sdr_pipe_run_fd(head.h, sink.h, _fd(inHandle), samples, maxVectors, -1 if outHandle is None else _fd(outHandle),
                                  0, C.byref(st)))
    return st


def runUdp(head, sink, sock, size, nVectors, outSock=None):
    """``runEffect $ udpSource sock size >-> head >-> ... >-> sink >-> udpSink``: one datagram of up to `size` BYTES per
    vector (NetworkStream.hs:33-35), nVectors datagrams; outSock must be a connected datagram socket (one vector per
    datagram, NetworkStream.hs:37-42) or None."""
    eb = head.in_dtype.itemsize
    st = L.IoStats()
    flags = L.SDR_IO_DATAGRAM_IN | (L.SDR_IO_DATAGRAM_OUT if outSock is not None else 0)
    L.check(L.lib.sdr_pipe_run_fd(head.h, sink.h, _fd(sock), size // eb, nVectors, -1 if outSock is None else _fd(outSock), flags,
                                  C.byref(st)))
    return st
