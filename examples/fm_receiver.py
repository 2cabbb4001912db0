#!/usr/bin/env python
"""FM broadcast receiver chain on a B200 -- the counterpart of the reference's examples/fm/fm.hs:30-41:

    runEffect $ sdrStream ... >-> P.map interleavedIQUnsignedByteToFloat >-> firDecimator deci 8192 >-> fmDemod
                >-> firResampler resp 8192 >-> firFilter filt 8192 >-> P.map (VG.map (* 0.2)) >-> pulseAudioSink

with the source and the sink replaced by file descriptors (a recording made with `rtl_sdr -s 1280000 -f <freq> iq.u8`,
raw 48 kHz float32 audio out), which is what SDR.Serialize.fromHandle / toHandle are for (Serialize.hs:78-83):

    python examples/fm_receiver.py iq.u8 audio.f32          # or `-` for stdin / stdout
    aplay -t raw -f FLOAT_LE -r 48000 -c 1 audio.f32

Every stage runs on the device; the convert / decimate / discriminate front end is one fused kernel, the stages hand
their vectors to each other inside HBM, and the whole run is one native loop (sdr_pipe_run_fd): 2 bytes per input sample
go up the PCIe link, 0.15 bytes come down.  The filters are windowed-sinc designs of the same shapes as the reference's
example (decimate by 8, resample 3/10, 64-tap symmetric audio filter); swap in your own taps freely.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sdr_b200  # noqa: E402


def windowed_sinc(n, cutoff, gain=1.0):
    """Hamming-windowed sinc centred at (n-1)/2 (the formulas of SDR.FilterDesign, FilterDesign.hs:33-68)"""
    k = np.arange(n, dtype=np.float64) - (n - 1) / 2.0
    h = np.sinc(2 * cutoff * k) * 2 * cutoff
    w = 0.54 - 0.46 * np.cos(2 * np.pi * np.arange(n) / (n - 1))
    return (gain * h * w).astype(np.float32)


def main():
    if len(sys.argv) != 3:
        sys.exit(__doc__)
    if not sdr_b200.has_cuda():
        sys.exit("no sm_100 device: sdr_b200 has no CPU path (use the reference's fast* constructors there)")
    fin = sys.stdin.buffer if sys.argv[1] == "-" else open(sys.argv[1], "rb")
    fout = sys.stdout.buffer if sys.argv[2] == "-" else open(sys.argv[2], "wb")

    deci = sdr_b200.cudaDecimatorC(8, windowed_sinc(128, 1 / 16), sizeMultiple=4)        # fm.hs:30  fastDecimatorC info 8 coeffsRFDecim
    resp = sdr_b200.cudaResamplerR(3, 10, windowed_sinc(90, 1 / 20, gain=3.0), sizeMultiple=8)   # fm.hs:31
    filt = sdr_b200.cudaFilterSymR(windowed_sinc(64, 1 / 4)[:32])                        # fm.hs:32  fastFilterSymR (half the taps)

    front = sdr_b200.pipeFmFrontEnd(deci, 8192)          # P.map convert >-> firDecimator deci 8192 >-> fmDemod, fused
    resampler = sdr_b200.pipeFirResampler(resp, 8192)
    audio = sdr_b200.pipeFirFilter(filt, 8192)
    volume = sdr_b200.pipeScale(0.2, front.ctx)
    front.connect(resampler).connect(audio).connect(volume)   # >-> on the device

    # vectors of 16384 bytes = 8192 IQ pairs, like the reference's 8192-sample buffers (fm.hs:24)
    try:
        st = sdr_b200.serialize.runHandles(front, volume, 16384, fin, fout)
    except sdr_b200.SdrError as e:
        # a recording whose tail is shorter than the decimator's 128 taps trips the reference's own `decimate 1`
        # assert (Filter.hs:586) -- after everything before it has been processed and written
        sys.exit(f"stopped at the end of the input: {e.msg}")
    print(f"{st.elements_in // 2} IQ samples in, {st.elements_out} audio samples out "
          f"(read {st.read_seconds:.3f} s, write {st.write_seconds:.3f} s)", file=sys.stderr)


if __name__ == "__main__":
    main()
