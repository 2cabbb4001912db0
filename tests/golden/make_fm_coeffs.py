#!/usr/bin/env python
"""Extracts the three coefficient lists of the reference's FM receiver example (examples/fm/Coeffs.hs:11-66, 76-110,
120-154) into tests/golden/fm_example_coeffs.npz.  Run in the build container, where /root/reference exists; the GPU box
only sees the committed .npz."""
import os
import re

import numpy as np

SRC = "/root/reference/examples/fm/Coeffs.hs"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fm_example_coeffs.npz")


def main():
    text = open(SRC).read()
    text = re.sub(r"\{-.*?-\}", "", text, flags=re.S)
    out = {}
    for m in re.finditer(r"^(coeffs\w+)\s*=\s*\[(.*?)\]", text, re.S | re.M):
        out[m.group(1)] = np.array([float(v) for v in re.findall(r"-?\d+\.\d+(?:[eE]-?\d+)?", m.group(2))], np.float32)
    assert [len(out[k]) for k in ("coeffsRFDecim", "coeffsAudioResampler", "coeffsAudioFilter")] == [51, 31, 32], {k: len(v) for k, v in out.items()}
    np.savez(OUT, **out)
    print({k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
