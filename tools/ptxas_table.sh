#!/bin/sh
# Registers / spills of the sweep kernels for a set of -D switches (no GPU needed):
#   tools/ptxas_table.sh -DPHYLO_FAST_BUILD -DPHYLO_PRETIP=0
cd "$(dirname "$0")/../phylostan_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I../../include \
    -Xptxas -v -c kernels.cu -o /tmp/kernels_$$.o "$@" 2>&1 | python3 -c '
import re, sys
txt = sys.stdin.read()
for m in re.finditer(r"Compiling entry function .(\S+?). for .sm_100a.\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", txt):
    name = m.group(1)
    k = re.search(r"(sweep_kernelI\w+?EE|stream_kernelI\w|contract_kernelI\w)", name)
    print(f"{k.group(1) if k else name[-40:]:60s} regs {m.group(5):>3s} stack {m.group(2):>4s} spill st/ld {m.group(3):>4s}/{m.group(4):>4s}")
'
rm -f /tmp/kernels_$$.o
