#!/usr/bin/env python
"""Summarise `ncu --set full` reports (read here with `ncu -i ... --page raw --csv`): one row per report with
the counters DESIGN.md / profiles/ quote.  Usage: python scripts/ncu_summary.py gpurun_out/r02_full_*.ncu-rep"""
import csv
import subprocess
import sys

WANT = [
    ("time", "gpu__time_duration.sum"), ("regs", "launch__registers_per_thread"),
    ("fmaheavy %", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    ("fma inst %", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
    ("alu %", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    ("fp64 %", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
    ("lsu %", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
    ("issue %", "sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
    ("warps/SM", "sm__warps_active.avg.per_cycle_active"),
    ("smem wavefronts %", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
    ("dram read", "dram__bytes_read.sum"), ("dram write", "dram__bytes_write.sum"),
    ("warp inst", "smsp__inst_executed.sum"),
]


def main(paths):
    for path in paths:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units = rows[0], rows[1]
        for vals in rows[2:]:
            get = lambda name: next((f"{vals[i]} {units[i]}".strip() for i, h in enumerate(hdr) if h == name), "n/a")
            print(f"## {get('Kernel Name')}  grid {get('Grid Size')} x block {get('Block Size')}   [{path}]")
            for label, name in WANT:
                print(f"  {label:18s} {get(name)}")
            stalls = []
            for i, h in enumerate(hdr):
                if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                    try:
                        v = float(vals[i])
                    except ValueError:
                        continue
                    key = h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]
                    if v >= 0.2 and key not in ("selected",) and not key.endswith("not_issued"):
                        stalls.append((v, key))
            print("  stalls (warps per issue slot): " + ", ".join(f"{k} {v:.2f}" for v, k in sorted(stalls, reverse=True)))


if __name__ == "__main__":
    main(sys.argv[1:])
