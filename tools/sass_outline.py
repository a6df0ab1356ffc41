#!/usr/bin/env python3
"""Static SASS outline of one kernel from `nvdisasm -gi <cubin>`: every instruction with the OUTERMOST source line
(the line inside the kernel body that the inlined helper was called from).
    nvdisasm -gi x.cubin > x.sass;  python tools/sass_outline.py x.sass <kernel-substring> [lo hi]
With lo/hi: opcode histogram of the instructions whose outer line falls in [lo, hi], per outer line."""
import collections
import re
import sys

text = open(sys.argv[1]).read().split('\n')
key = sys.argv[2]
lo = int(sys.argv[3]) if len(sys.argv) > 3 else 0
hi = int(sys.argv[4]) if len(sys.argv) > 4 else 10 ** 9
inside, outer, inner, fresh = False, None, None, True
rows = []
for ln in text:
    if ln.startswith('\t.section'):
        inside = key in ln and '.text.' in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File ".*?", line (\d+)( inlined at)?', ln)
    if m:
        if fresh:
            inner = int(m.group(1))
            fresh = False
        outer = int(m.group(1))
        continue
    m = re.match(r'\s+/\*([0-9a-f]+)\*/\s+(.*?);', ln)
    if m:
        rows.append((int(m.group(1), 16), outer, inner, m.group(2).strip()))
        fresh = True
    elif re.match(r'\.L_x_\d+:', ln.strip()):
        rows.append((None, None, None, ln.strip()))
if len(sys.argv) <= 3:
    for a, o, i, t in rows:
        print(f"{a if a is None else hex(a)}\t{o}\t{i}\t{t}")
else:
    per = collections.defaultdict(collections.Counter)
    for a, o, i, t in rows:
        if a is None or o is None or not (lo <= o <= hi):
            continue
        op = t.split()
        op = op[1] if op[0].startswith('@') else op[0]
        op = 'MOV' if op.startswith('IMAD.MOV') else op.split('.')[0]
        per[o][op] += 1
    tot = collections.Counter()
    for o in sorted(per):
        tot += per[o]
        print(o, sum(per[o].values()), dict(per[o].most_common(8)))
    print('TOTAL', sum(tot.values()), tot.most_common(30))
