#!/usr/bin/env python
"""Turn the outputs of tools/profile_round.sh (gpurun_out/<tag>_*) into the committed summaries under profiles/:
  profiles/<tag>_train_step_b8_launches.md   per-kernel shares of one eager training step (ncu launch list)
  profiles/<tag>_ncu_kernels.md              key `ncu --set full` metrics of the top kernels
  profiles/<tag>_gemm_dram.json              DRAM bytes / duration of every GEMM launch of one step (roofline.traffic)
usage: tools/make_profiles.py TAG"""
import collections
import csv
import glob
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")

# 1. launch list
out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "summarize_launches.py"), os.path.join(G, f"{tag}_train_launches.csv"), "2"],
                     capture_output=True, text=True).stdout
open(os.path.join(P, f"{tag}_train_step_b8_launches.md"), "w").write(
    "One eager b=8 training step (tests/probe_train_once.py, CUDA graph off so that ncu sees every launch), "
    "`ncu --metrics gpu__time_duration.sum --clock-control none`.\n\n" + out)

# 2. full captures
KEYS = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("lts__t_bytes.sum", "L2 bytes"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor instr"), ("smsp__inst_executed.sum", "warp instr"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"), ("launch__registers_per_thread", "regs/thread"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("sm__cycles_elapsed.max", "SM cycles"), ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts")]
rows_md = []
for f in sorted(glob.glob(os.path.join(G, f"{tag}_full_*.raw.csv"))):
    rows = list(csv.reader(open(f)))
    if len(rows) < 3:
        continue
    d = {h: (v, u) for h, u, v in zip(rows[0], rows[1], rows[2])}
    name = os.path.basename(f)[len(tag) + 6:-8]

    def num(k):
        v, u = d.get(k, ("", ""))
        try:
            x = float(v.replace(",", ""))
        except ValueError:
            return None, u
        return x, u
    dur, du = num("gpu__time_duration.sum")
    dur_s = dur * {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}.get(du, 1e-6) if dur else None
    tot = 0.0
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        x, u = num(k)
        if x is not None:
            tot += x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    line = [f"### {name}", "", f"kernel: `{d.get('Kernel Name', ('?',))[0][:90]}`", "", "| metric | value |", "|---|---|"]
    for k, label in KEYS:
        if k in d:
            line.append(f"| {label} (`{k}`) | {d[k][0]} {d[k][1]} |")
    if dur_s:
        line.append(f"| **DRAM traffic / duration** | {tot / 1e6:.1f} MB / {dur_s * 1e6:.1f} us = {tot / dur_s / 1e9:.0f} GB/s |")
    rows_md.append("\n".join(line))
open(os.path.join(P, f"{tag}_ncu_kernels.md"), "w").write(
    f"`ncu --set full --clock-control none --import-source on`, one launch per kernel (tools/profile_round.sh {tag}); the .ncu-rep files stay in "
    "gpurun_out/ (scratch).  cgemm captures come from tests/probe_cgemm.py (bf = 88 frames), the others from the first matching launch of a "
    "b=8 training step (level 0: 96x96x11).\n\n" + "\n\n".join(rows_md) + "\n")

# 3. DRAM traffic of every GEMM launch
f = os.path.join(G, f"{tag}_gemm_dram.csv")
if os.path.exists(f):
    rows = list(csv.reader(l for l in open(f) if l.startswith('"')))
    idx = {h: i for i, h in enumerate(rows[0])}
    per = collections.OrderedDict()
    for r in rows[1:]:
        k = r[idx["ID"]]
        e = per.setdefault(k, {"kernel": "cgemm" if "cgemm" in r[idx["Kernel Name"]] else "wgrad"})
        v = float(r[idx["Metric Value"]].replace(",", ""))
        u = r[idx["Metric Unit"]]
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3}.get(u, 1)
        e[r[idx["Metric Name"]]] = v
    agg = {}
    for k in ("cgemm", "wgrad"):
        es = [e for e in per.values() if e["kernel"] == k]
        # the probe runs ONE step after model construction: every launch belongs to it
        agg[k] = {"launches": len(es), "dram_bytes_total": sum(e.get("dram__bytes_read.sum", 0) + e.get("dram__bytes_write.sum", 0) for e in es),
                  "duration_s_total_cold": sum(e.get("gpu__time_duration.sum", 0) for e in es)}
        agg[k]["dram_bytes_per_launch"] = agg[k]["dram_bytes_total"] / max(1, len(es))
    json.dump(agg, open(os.path.join(P, f"{tag}_gemm_dram.json"), "w"), indent=1)
    print(agg)
