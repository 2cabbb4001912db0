#!/usr/bin/env python
"""Summarise an .ncu-rep (read here with `ncu -i ... --page raw --csv`) into the handful of numbers DESIGN.md quotes."""
import csv, subprocess, sys
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]
for path in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("==", path, "::", r[hdr.index("Kernel Name")][:110])
        for k in KEYS:
            if k in hdr:
                print(f"  {k:75s} {r[hdr.index(k)]} {units[hdr.index(k)]}")
        st = [(h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", ""), float(r[i])) for i, h in enumerate(hdr)
              if "smsp__average_warps_issue_stalled" in h and "per_issue_active" in h]
        print("  stall reasons (warps per issue):", ", ".join(f"{n}={v:.2f}" for n, v in sorted(st, key=lambda t: -t[1])[:7]))
