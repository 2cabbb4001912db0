"""Generate golden vectors from the UNMODIFIED reference C (oracle/_ref/libsdrref.so).

Run in the build container (where /root/reference exists):   python tests/golden/make_golden.py
Writes tests/golden/ref_c_golden.npz: for every one of the reference's 42 DSP symbols (all of c_sources except
cpuid/cpuid_extended) the output on seeded inputs.  Inputs are regenerated from the seed by the tests
(`golden_inputs`), so only outputs are stored.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from oracle import pipes  # noqa: E402

SEED = 0x5D2B200
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_c_golden.npz")

# (size, half taps, factor) -- generator shapes of the reference's own tests (TestSuite.hs:55-57), kept small
CASES = [(1024, 32, 1), (2048, 64, 8), (1024, 32, 3)]
# (size, taps, L, M, starting group)
RCASES = [(1024, 90, 3, 10, 0), (2048, 77, 3, 7, 2), (1024, 64, 5, 11, 3)]


def golden_inputs():
    """Seeded inputs shared by the generator and the tests; values in (-10, 10) like TestSuite.hs:62."""
    rng = np.random.default_rng(SEED)
    d = {}
    for ci, (size, half, factor) in enumerate(CASES):
        d[f"c{ci}_half"] = rng.uniform(-10, 10, half).astype(np.float32)
        d[f"c{ci}_xr"] = rng.uniform(-10, 10, size).astype(np.float32)
        d[f"c{ci}_xc"] = (rng.uniform(-10, 10, size) + 1j * rng.uniform(-10, 10, size)).astype(np.complex64)
    for ri, (size, taps, L, M, g) in enumerate(RCASES):
        d[f"r{ri}_taps"] = rng.uniform(-10, 10, taps).astype(np.float32)
        d[f"r{ri}_xr"] = rng.uniform(-10, 10, size).astype(np.float32)
        d[f"r{ri}_xc"] = (rng.uniform(-10, 10, size) + 1j * rng.uniform(-10, 10, size)).astype(np.complex64)
    d["u8"] = rng.integers(0, 256, 4096, dtype=np.uint8)
    d["i16"] = rng.integers(-2048, 2048, 4096).astype(np.int16)
    d["tx"] = np.concatenate([rng.uniform(-1, 1, 4000), [-1.0, 0.0, 0.99999, -0.5, 1.0, -1.5, 2.5, 0.25]]
                             ).astype(np.float32)
    d["sc"] = rng.uniform(-10, 10, 4096).astype(np.float32)
    d["dc"] = rng.uniform(-1, 1, 4096).astype(np.float32)
    return d


def coeff_layouts(half):
    full = np.concatenate([half, half[::-1]])
    return {"plain": full, "dup": pipes.duplicate(full), "half": half}


# reference symbol -> coefficient layout it expects
LAYOUT_R = {"RR": "plain", "SSERR": "plain", "AVXRR": "plain", "SSESymmetricRR": "half", "AVXSymmetricRR": "half"}
LAYOUT_C = {"RC": "plain", "SSERC": "dup", "AVXRC": "dup", "SSERC2": "plain", "AVXRC2": "plain",
            "SSESymmetricRC": "half", "AVXSymmetricRC": "half"}


def main():
    ref = oracle.ref()
    if ref is None:
        raise SystemExit("oracle/_ref/libsdrref.so missing: run `make -C oracle` where /root/reference exists")
    inp = golden_inputs()
    out = {}
    for ci, (size, half_n, factor) in enumerate(CASES):
        lay = coeff_layouts(inp[f"c{ci}_half"])
        T = 2 * half_n
        numf = size - T + 1
        numd = (size - T) // factor + 1
        for suf, l in LAYOUT_R.items():
            out[f"c{ci}_filter{suf}"] = ref.filter("filter" + suf, numf, lay[l], inp[f"c{ci}_xr"])
            out[f"c{ci}_decimate{suf}"] = ref.decimate("decimate" + suf, numd, factor, lay[l], inp[f"c{ci}_xr"])
        for suf, l in LAYOUT_C.items():
            out[f"c{ci}_filter{suf}"] = ref.filter("filter" + suf, numf, lay[l], inp[f"c{ci}_xc"])
            out[f"c{ci}_decimate{suf}"] = ref.decimate("decimate" + suf, numd, factor, lay[l], inp[f"c{ci}_xc"])
    for ri, (size, taps_n, L, M, g0) in enumerate(RCASES):
        taps = inp[f"r{ri}_taps"]
        num = (size * L - pipes.round_up(taps_n, 8 * L)) // M + 1 - 8
        off = L - 1 - ((L + g0 * M - 1) % L)
        out[f"r{ri}_resampleRR"] = ref.resample_legacy(num, L, M, off, taps, inp[f"r{ri}_xr"])
        for name, sm in (("resample2RR", 1), ("resampleSSERR", 4), ("resampleAVXRR", 8)):
            nc, inc, groups = pipes.prepare_coeffs(sm, L, M, taps)
            y, g = ref.resample(name, num, nc, g0, inc, groups, inp[f"r{ri}_xr"])
            out[f"r{ri}_{name}"] = y
            out[f"r{ri}_{name}_group"] = np.int32(g)
        for name, sm in (("resample2RC", 1), ("resampleSSERC", 4), ("resampleAVXRC", 8)):
            nc, inc, groups = pipes.prepare_coeffs(sm, L, M, taps)
            y, g = ref.resample(name, num, nc, g0, inc, groups, inp[f"r{ri}_xc"])
            out[f"r{ri}_{name}"] = y
            out[f"r{ri}_{name}_group"] = np.int32(g)
    for n in ("convertC", "convertCSSE", "convertCAVX"):
        out[n] = ref.convert_u8(n, inp["u8"])
    for n in ("convertCBladeRF", "convertCSSEBladeRF", "convertCAVXBladeRF"):
        out[n] = ref.convert_i16(n, inp["i16"])
    out["convertBladeRFTransmit"] = ref.convert_tx(inp["tx"])
    for n in ("scale", "scaleSSE", "scaleAVX"):
        out[n] = ref.scale(n, np.float32(0.2), inp["sc"])
    y, fs, fo = ref.dc_blocker(inp["dc"], 0.25, -0.125)
    out["dcBlocker"] = y
    out["dcBlocker_final"] = np.array([fs, fo], np.float32)
    np.savez_compressed(OUT, **out)
    print(f"wrote {OUT}: {len(out)} arrays, {os.path.getsize(OUT)/1024:.0f} KiB")


if __name__ == "__main__":
    main()
