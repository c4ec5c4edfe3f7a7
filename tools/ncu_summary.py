#!/usr/bin/env python
"""Compact per-kernel summary of an .ncu-rep (read on the CPU box): python tools/ncu_summary.py rep.ncu-rep [out.md]

Pulls the handful of metrics the roofline discussion needs out of `ncu --page raw --csv` so the summary can be
committed under profiles/ (the .ncu-rep itself stays in gpurun_out/, which is scratch).
"""
from __future__ import annotations

import csv
import io
import subprocess
import sys

EXACT = [
    "gpu__time_duration.sum",
    "sm__cycles_elapsed.max",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum",
    "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor.sum",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.sum",
    "smsp__inst_executed.avg.per_cycle_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "launch__registers_per_thread",
    "launch__grid_size",
    "launch__block_size",
    "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers",
    "launch__waves_per_multiprocessor",
]


def main() -> None:
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    extra = [h for h in hdr if ("stall" in h and h.endswith("_per_warp_active.pct")) or h.startswith("smsp__average_warps_issue_stalled")
             and h.endswith("_per_issue_active.ratio")]
    lines = [f"# ncu summary of `{rep}`", ""]
    for r in body:
        lines.append(f"## {r[col['Kernel Name']][:140]}")
        lines.append(f"grid {r[col['Grid Size']]} block {r[col['Block Size']]}")
        lines.append("")
        lines.append("| metric | value | unit |")
        lines.append("|---|---|---|")
        for m in EXACT:
            if m in col and r[col[m]] not in ("", "n/a"):
                lines.append(f"| {m} | {r[col[m]]} | {units[col[m]]} |")
        stalls = []
        for m in extra:
            try:
                stalls.append((float(r[col[m]].replace(",", "")), m))
            except ValueError:
                pass
        for v, m in sorted(stalls, reverse=True)[:8]:
            lines.append(f"| {m} | {v:.3f} | {units[col[m]]} |")
        lines.append("")
    text = "\n".join(lines)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text + "\n")
    print(text)


if __name__ == "__main__":
    main()
