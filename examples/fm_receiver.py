#!/usr/bin/env python
"""FM broadcast receiver chain on a B200 -- the counterpart of the reference's examples/fm/fm.hs:30-41:

    runEffect $ sdrStream ... >-> P.map interleavedIQUnsignedByteToFloat >-> firDecimator deci 8192 >-> fmDemod
                >-> firResampler resp 8192 >-> firFilter filt 8192 >-> P.map (VG.map (* 0.2)) >-> pulseAudioSink

with the source and the sink replaced by file descriptors (a recording made with `rtl_sdr -s 1280000 -f <freq> iq.u8`,
raw 48 kHz float32 audio out), which is what SDR.Serialize.fromHandle / toHandle are for (Serialize.hs:78-83):

    python examples/fm_receiver.py iq.u8 audio.f32          # or `-` for stdin / stdout
    aplay -t raw -f FLOAT_LE -r 48000 -c 1 audio.f32

The chain is TWO fused kernels on the device -- convert / decimate / discriminate, and resample / filter / volume -- that
hand their vectors to each other inside HBM, and the whole run is one native loop (sdr_pipe_run_fd): 2 bytes per input
sample go up the PCIe link, 0.15 bytes come down.  `--reference-coeffs` uses the reference example's own coefficient sets
(examples/fm/Coeffs.hs, from tests/golden/fm_example_coeffs.npz: 51-tap RF decimator, 31-tap resampler, 32 half-taps);
otherwise the filters are windowed-sinc designs of the BASELINE shapes (128 / 90 / 64 taps).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sdr_b200  # noqa: E402


windowed_sinc = sdr_b200.windowed_sinc_taps   # the formulas of SDR.FilterDesign (FilterDesign.hs:33-68)


def main():
    ref_coeffs = "--reference-coeffs" in sys.argv
    if ref_coeffs:
        sys.argv.remove("--reference-coeffs")
    if len(sys.argv) != 3:
        sys.exit(__doc__)
    if not sdr_b200.has_cuda():
        sys.exit("no sm_100 device: sdr_b200 has no CPU path (use the reference's fast* constructors there)")
    fin = sys.stdin.buffer if sys.argv[1] == "-" else open(sys.argv[1], "rb")
    fout = sys.stdout.buffer if sys.argv[2] == "-" else open(sys.argv[2], "wb")

    if ref_coeffs:
        fm = np.load(os.path.join(ROOT, "tests", "golden", "fm_example_coeffs.npz"))
        c_rf, c_rs, c_au = fm["coeffsRFDecim"], fm["coeffsAudioResampler"], fm["coeffsAudioFilter"]
    else:
        c_rf, c_rs, c_au = windowed_sinc(128, 1 / 16), windowed_sinc(90, 1 / 20, gain=3.0), windowed_sinc(64, 1 / 4)[:32]
    deci = sdr_b200.cudaDecimatorC(8, c_rf, sizeMultiple=4)        # fm.hs:30  fastDecimatorC info 8 coeffsRFDecim
    resp = sdr_b200.cudaResamplerR(3, 10, c_rs, sizeMultiple=8)    # fm.hs:31  fastResamplerR info 3 10 coeffsAudioResampler
    filt = sdr_b200.cudaFilterSymR(c_au)                           # fm.hs:32  fastFilterSymR info coeffsAudioFilter (half the taps)

    front = sdr_b200.pipeFmFrontEnd(deci, 8192)                    # P.map convert >-> firDecimator deci 8192 >-> fmDemod, fused
    low = sdr_b200.pipeFmLowRate(resp, 8192, filt, 8192, 0.2)      # firResampler resp 8192 >-> firFilter filt 8192 >-> P.map (* 0.2), fused
    front.connect(low)                                             # >-> on the device

    # vectors of 16384 bytes = 8192 IQ pairs, like the reference's 8192-sample buffers (fm.hs:24)
    try:
        st = sdr_b200.serialize.runHandles(front, low, 16384, fin, fout)
    except sdr_b200.SdrError as e:
        # a recording whose tail is shorter than the decimator's 128 taps trips the reference's own `decimate 1`
        # assert (Filter.hs:586) -- after everything before it has been processed and written
        sys.exit(f"stopped at the end of the input: {e.msg}")
    print(f"{st.elements_in // 2} IQ samples in, {st.elements_out} audio samples out "
          f"(read {st.read_seconds:.3f} s, write {st.write_seconds:.3f} s)", file=sys.stderr)


if __name__ == "__main__":
    main()
