#!/usr/bin/env python
"""Sweep of the dcBlocker speculation parameters (chunk length, warm-up lengths) on device-resident noise; prints one
JSON line per setting.  Measurement aid for the automatic choice in sdr_b200/csrc/kernels_dc.cu."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sdr_b200  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def main():
    ctx = sdr_b200.default_context()
    for log2n in [int(a) for a in sys.argv[1:]] or [28]:
        n = 1 << log2n
        x, y, fin = ctx.alloc(4 * n), ctx.alloc(4 * n), ctx.alloc(8)
        ctx.synth_noise(x, n)
        settings = ((0, -1, -1), (4096, -1, -1), (8192, -1, -1)) if os.environ.get("DC_SWEEP_SHORT") else None
        for chunk, k1, k2 in settings or ((0, -1, -1), (2048, -1, -1), (3072, -1, -1), (4096, -1, -1), (6144, -1, -1), (8192, -1, -1),
                              (12288, -1, -1), (16384, -1, -1), (8192, 4096, -1), (8192, 8192, -1), (8192, -1, 3072), (4096, -1, 3072)):
            ctx.dc_tuning(chunk, k1, k2, -1)
            for _ in range(2):
                ctx.dc_blocker(x.ptr, y.ptr, n, fin.ptr)
            ctx.sync()
            st0, _ = ctx.dc_stats()
            e0, e1 = sdr_b200.Event(ctx), sdr_b200.Event(ctx)
            e0.record()
            for _ in range(5):
                ctx.dc_blocker(x.ptr, y.ptr, n, fin.ptr)
            e1.record()
            ms = e0.elapsed_ms(e1) / 5
            st1, par = ctx.dc_stats()
            gs = n / (ms * 1e-3) / 1e9
            print(json.dumps({"log2n": log2n, "chunk": chunk, "k1": k1, "k2": k2, "ms": round(ms, 4), "Gsamples_per_s": round(gs, 1),
                              "hbm_frac": round(gs * 8 / PEAK, 3), "chunks_per_call": (st1[1] - st0[1]) // 5,
                              "repaired_per_call": (st1[2] - st0[2]) / 5, "parallel": par,
                              "mode": os.environ.get("SDR_B200_DC_MODE", "default")}), flush=True)
        ctx.dc_tuning()
        for b in (x, y, fin):
            b.free()


if __name__ == "__main__":
    main()
