#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, no GPU): key metrics per captured launch.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [out.md]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
keys = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg"]
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
pipes = [h for h in hdr if h.startswith("sm__inst_executed_pipe_") and h.endswith(".sum")]
out = []
for r in data:
    out.append("## " + r[hdr.index("Kernel Name")][:90])
    for k in keys[1:]:
        if k in hdr:
            i = hdr.index(k)
            out.append(f"- {k}: {r[i]} {units[i]}")
    st = sorted(((float(r[hdr.index(s)] or 0), s) for s in stall), reverse=True)[:7]
    out.append("- top stalls (warps per issue): " + ", ".join(f"{s.split('stalled_')[1].split('_per_issue')[0]}={v:.2f}" for v, s in st))
    pp = sorted(((float(r[hdr.index(s)] or 0), s) for s in pipes), reverse=True)[:9]
    out.append("- inst by pipe: " + ", ".join(f"{s.split('pipe_')[1].split('.sum')[0]}={v:.3g}" for v, s in pp))
    out.append("")
text = "\n".join(out)
print(text)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text + "\n")
