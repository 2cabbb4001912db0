#!/usr/bin/env python
"""One section of profiles/r02_scaling_8gpu.txt from the bench lines of a scaling run: python tools/scaling_table.py gpurun_out/r2_final8 "title" """
import json, sys
prefix, title = sys.argv[1], sys.argv[2]
lines = {}
for n in (1, 2, 4, 8):
    try:
        lines[n] = json.loads(open(f"{prefix}_n{n}.json").read().strip().splitlines()[-1])
    except Exception as e:
        print(f"# N={n}: no line ({e})")
print(f"\n## {title}")
print("N | value Msamples/s | x vs N=1 | efficiency | ms/pass (max rank) | per-rank ms/pass | roofline frac (max rank) | sustained Msamples/s (frac) | "
      "halo ms/pass | NCCL-transport ms/pass | e2e f32 | plain H2D GB/s per GPU | e2e_u8 | checksum | clocks MHz / reasons")
base = lines[1]["value"] if 1 in lines else None
for n, d in lines.items():
    e2e, u8, sus = d.get("e2e") or {}, d.get("e2e_u8") or {}, d.get("sustained") or {}
    print(" | ".join(str(v) for v in [
        n, round(d["value"]), round(d["value"] / base, 3) if base else None, round(d["value"] / base / n, 3) if base else None,
        round(d.get("ms_per_pass", 0), 5), [round(x, 5) for x in d.get("ms_per_pass_by_rank", [])], round(d["roofline"]["frac"], 3),
        f"{round(sus.get('value', 0))} ({round(sus.get('roofline_frac', 0), 3)})", d.get("halo_ms_per_pass") and round(d["halo_ms_per_pass"], 5),
        d.get("nccl_halo_ms_per_pass") and round(d["nccl_halo_ms_per_pass"], 5), e2e.get("value") and round(e2e["value"]),
        e2e.get("pcie_h2d_GBps_plain_memcpy") and round(e2e["pcie_h2d_GBps_plain_memcpy"], 1), u8.get("value") and round(u8["value"]),
        d["config"].get("output_checksum"), f"{d['clocks'].get('sm_mhz')} {d['clocks'].get('reasons')}"]))
for n, d in lines.items():
    print(f"\n### raw line N={n}")
    print(json.dumps(d))
