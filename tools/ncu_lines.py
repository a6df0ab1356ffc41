#!/usr/bin/env python3
"""Per-CUDA-line and per-opcode dynamic instruction counts from `ncu --page source --print-source cuda,sass --csv`.
    python tools/ncu_lines.py src_cuda.csv <warp_steps>"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))[3:]
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0


def I(x):
    try:
        return int(x)
    except ValueError:
        return 0


lines, cur = [], None
byop, samp = collections.Counter(), collections.Counter()
for r in rows:
    if r[0] and r[0].isdigit():
        cur = [int(r[0]), r[1], 0, 0, 0]
        lines.append(cur)
    elif cur is not None and len(r) > 7 and r[3].split():
        op = r[3].split()
        o = (op[1] if op[0].startswith('@') and len(op) > 1 else op[0]).split('.')[0]
        n, s = I(r[7]), I(r[6])
        cur[2] += n
        cur[3] += s
        if o in ('DFMA', 'DMUL', 'DADD'):
            cur[4] += n
        byop[o] += n
        samp[o] += s
tot = sum(byop.values())
ts = sum(samp.values())
print(f"total {tot / steps:.1f} instr per warp-step; fp64 {sum(byop[o] for o in ('DFMA','DMUL','DADD')) / steps:.1f}")
print("opcode mix:", ", ".join(f"{o}={n / steps:.0f}({100 * samp[o] / ts:.0f}%)" for o, n in byop.most_common(24)))
lines.sort(key=lambda x: -x[3])
for ln, src, n, s, f in lines[:32]:
    print(f"{ln:4d} instr {n / steps:7.1f} fp64 {f / steps:6.1f} samples {100 * s / ts:5.1f}%  {src[:105]}")
