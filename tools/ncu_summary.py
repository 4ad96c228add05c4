#!/usr/bin/env python3
"""Summarise `ncu --set full` captures (gpurun_out/*.ncu-rep) into small text files under profiles/.

    python tools/ncu_summary.py gpurun_out/r01a_prof_imdct.ncu-rep [...]

Writes profiles/<name>.summary.txt with the metrics the roofline discussion uses (duration, DRAM bytes, issue
rate, pipe utilisation, occupancy, stall reasons) and prints them. Reads with `ncu -i … --page raw --csv`,
which needs no GPU.
"""
from __future__ import annotations

import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = (
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__t_bytes.sum",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max", "sm__cycles_elapsed.max.per_second",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmalite.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
)


def summarise(path: str) -> str:
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv", "--print-units", "base"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = [f"# {os.path.basename(path)} — ncu --set full --clock-control none (one launch per row; cold cache, serialised)"]
    for row in rows[2:]:
        col = dict(zip(hdr, row))
        out.append(f"kernel: {col.get('Kernel Name')}  grid {col.get('Grid Size')} block {col.get('Block Size')}")
        for i, h in enumerate(hdr):
            if h in KEEP or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
                out.append(f"  {h} = {row[i]} {units[i]}")
        rd, wr = col.get("dram__bytes_read.sum"), col.get("dram__bytes_write.sum")
        if rd and wr:
            out.append(f"  traffic (dram read + write) = {float(rd) + float(wr):.6f} {units[hdr.index('dram__bytes_read.sum')]}")
    return "\n".join(out) + "\n"


def main():
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    for p in sys.argv[1:]:
        text = summarise(p)
        name = os.path.splitext(os.path.basename(p))[0]
        with open(os.path.join(ROOT, "profiles", name + ".summary.txt"), "w") as f:
            f.write(text)
        print(text)


if __name__ == "__main__":
    main()
