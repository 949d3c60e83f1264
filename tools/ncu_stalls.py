#!/usr/bin/env python
"""Print the warp-stall sample breakdown and a few headline metrics of one `ncu --set full` raw-page CSV (ncu -i X.ncu-rep --page raw --csv)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
d = {h: (v, u) for h, u, v in zip(rows[0], rows[1], rows[2])}
st = []
for k, (v, u) in d.items():
    if "pcsamp_warps_issue_stalled" in k and "not_issued" not in k:
        try:
            x = float(v.replace(",", ""))
        except ValueError:
            continue
        if x > 0:
            st.append((x, k.replace("smsp__pcsamp_warps_issue_stalled_", "")))
tot = sum(x for x, _ in st) or 1.0
for x, k in sorted(st, reverse=True):
    print(f"{k:32s} {x:10.0f}  {100 * x / tot:5.1f} %")
for k in ["gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
          "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
          "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
          "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem"]:
    print(k, d.get(k))
