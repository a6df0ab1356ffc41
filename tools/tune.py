#!/usr/bin/env python3
"""Sweep the tile shapes of the sweep kernel on the BASELINE config-3 problem (GPU box only).
    python tools/tune.py [B] [L]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from phylostan_b200 import likelihood as lk, synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
L = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
S = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
t0 = time.time()
prob = synth.make_problem(S, L, 4, structured=False)
draws = synth.make_draws(prob, B)
print(f"problem {S}x{L} built in {time.time() - t0:.1f}s", flush=True)
lik = lk.TreeLikelihood(prob.peel, prob.tipmask, prob.weights, model="GTR", categories=4)
PREC = int(os.environ.get("PHYLO_PREC", "64"))
lik.set_precision(PREC)
lik.set_stack_slots(int(os.environ.get("PHYLO_SLOTS", "0")))
print("precision", PREC, flush=True)
lik.upload(*draws)
alg = (32.0 * L * 4 * (5 * S - 9) + 2.0 * S * L + 8.0 * L) * B
ref = None
for grad in (True, False):
    for K, PB in ((1, 1), (1, 2), (1, 4), (2, 1), (2, 2), (4, 1)):
        lik.set_tiling(K, PB)
        lik.set_timing(False)
        for _ in range(2):
            lik.run(B, grad)
        lik.sync()
        lik.set_timing(True)
        ms = []
        for _ in range(3):
            lik.run(B, grad)
            ms.append(lik.get_timing())
        out = lik.download(B)
        if ref is None:
            ref = out
        err = np.abs(out[:, 0] - ref[:, 0]).max() / np.abs(ref[:, 0]).max()
        sw = min(m["sweep_ms"] for m in ms)
        info = lik.info()
        print(f"grad={int(grad)} K={K} PB={PB} NT={info['threads_per_cta']} grid={info['grid']} smem={info['smem_bytes']} "
              f"sweep={sw:.2f} ms pmat={ms[-1]['pmat_ms']:.3f} contract={ms[-1]['contract_ms']:.3f} "
              f"-> {B / sw * 1e3:.1f} evals/s, B_vg-roofline frac {alg / (sw * 1e-3) / 6537.3e9:.3f}  relerr {err:.1e}",
              flush=True)
