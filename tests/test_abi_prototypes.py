"""Layer 1 of include/sdr_b200.h against the reference's own C prototypes (the symbols FilterInternal.hs / Util.hs bind with
`foreign import ccall`): every replacement takes the same arguments in the same order with the same types, and the one
family whose return value carries information (resample*: the next group, resample.c:34-142) returns it the same way.
`void` becomes an `int` status -- the only licensed difference.

The reference prototypes are parsed from /root/reference/c_sources when that tree is present (this container) and
pinned in tests/golden/ref_prototypes.json, which is what a box without the reference tree checks against."""
import json
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/c_sources"
FIXTURE = os.path.join(ROOT, "tests", "golden", "ref_prototypes.json")

# reference symbol -> its replacement
MAP = {}
for isa in ("", "SSE", "AVX"):
    MAP[f"filter{isa}RR"] = "filterCudaRR"
    MAP[f"decimate{isa}RR"] = "decimateCudaRR"
for isa in ("SSE", "AVX"):
    MAP[f"filter{isa}SymmetricRR"] = "filterCudaSymmetricRR"
    MAP[f"filter{isa}SymmetricRC"] = "filterCudaSymmetricRC"
    MAP[f"decimate{isa}SymmetricRR"] = "decimateCudaSymmetricRR"
    MAP[f"decimate{isa}SymmetricRC"] = "decimateCudaSymmetricRC"
    MAP[f"filter{isa}RC"] = "filterCudaRCDup"          # duplicated-coefficient forms
    MAP[f"decimate{isa}RC"] = "decimateCudaRCDup"
    MAP[f"filter{isa}RC2"] = "filterCudaRC"
    MAP[f"decimate{isa}RC2"] = "decimateCudaRC"
    MAP[f"resample{isa}RR"] = "resampleCudaRR"
    MAP[f"resample{isa}RC"] = "resampleCudaRC"
    MAP[f"scale{isa}"] = "scaleCuda"
    MAP[f"convertC{isa}"] = "convertCuda"
    MAP[f"convertC{isa}BladeRF"] = "convertCudaBladeRF"
MAP.update({"filterRC": "filterCudaRC", "decimateRC": "decimateCudaRC", "resample2RR": "resampleCudaRR",
            "resample2RC": "resampleCudaRC", "resampleRR": "resampleCudaLegacyRR", "scale": "scaleCuda",
            "convertC": "convertCuda", "convertCBladeRF": "convertCudaBladeRF",
            "convertBladeRFTransmit": "convertCudaBladeRFTransmit", "dcBlocker": "dcBlockerCuda"})
NOT_REPLACED = {"cpuid", "cpuid_extended"}   # x86 feature probe; its role is played by sdr_has_cuda


def _norm_args(arglist):
    out = []
    for a in arglist.split(","):
        a = re.sub(r"\bconst\b", " ", a)
        a = re.sub(r"\s+", " ", a).strip()
        m = re.match(r"^(.*?)(\w+)$", a)          # drop the parameter name
        t = (m.group(1) if m else a).replace(" ", "")
        out.append(t)
    return out


def parse_reference():
    protos = {}
    for fn in sorted(os.listdir(REF)):
        if not fn.endswith(".c"):
            continue
        for m in re.finditer(r"^(void|int)\s+(\w+)\s*\(([^)]*)\)\s*\{", open(os.path.join(REF, fn)).read(), re.M):
            protos[m.group(2)] = {"ret": m.group(1), "args": _norm_args(m.group(3)), "file": fn}
    return protos


def parse_header():
    text = open(os.path.join(ROOT, "include", "sdr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(int|void)\s+(\w+)\s*\(([^;{]*)\)\s*;", text):
        protos[m.group(2)] = {"ret": m.group(1), "args": _norm_args(re.sub(r"\s+", " ", m.group(3)))}
    return protos


def reference_prototypes():
    if os.path.isdir(REF):
        return parse_reference()
    return json.load(open(FIXTURE))


def test_fixture_pins_the_reference_prototypes():
    if not os.path.isdir(REF):
        pytest.skip("no reference tree on this box: the fixture is the pin")
    assert json.load(open(FIXTURE)) == parse_reference(), "regenerate with tests/golden/make_golden.py"


def test_every_reference_symbol_has_a_same_signature_replacement():
    ref = reference_prototypes()
    ours = parse_header()
    assert len(ref) == 44 and set(ref) - set(MAP) == NOT_REPLACED, sorted(set(ref) - set(MAP))
    for name, p in sorted(ref.items()):
        if name in NOT_REPLACED:
            continue
        q = ours[MAP[name]]
        assert q["args"] == p["args"], (name, MAP[name], p["args"], q["args"])
        assert q["ret"] == "int", (name, q["ret"])   # void -> status; int (resample*: next group) stays int
