"""Single-draw latency on mid-size problems (one chain of NUTS / one Stan gradient): the regime where the
sweep leaves most schedulers with one warp.   [PHYLO_K=1|2|4] python tools/midsize_latency.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from phylostan_b200 import likelihood as lk, synth  # noqa: E402

for S, L, B in ((100, 1000, 1), (200, 3000, 1), (500, 5000, 1), (500, 20000, 1), (1000, 10000, 1), (200, 3000, 4)):
    prob = synth.make_problem(S, L, 4, structured=False)
    draws = synth.make_draws(prob, B)
    with lk.TreeLikelihood(prob.peel, prob.tipmask, prob.weights, model="GTR", categories=4) as lik:
        lik.set_tiling(int(os.environ.get("PHYLO_K", "0")), 0)
        for _ in range(5):
            lik.value_grad(*draws)
        t0 = time.perf_counter()
        n = 30
        for _ in range(n):
            lik.value_grad(*draws)
        tg = (time.perf_counter() - t0) / n
        t0 = time.perf_counter()
        for _ in range(n):
            lik.loglik(*draws)
        tv = (time.perf_counter() - t0) / n
        i = lik.info()
        print(f"S={S:5d} L={L:6d} B={B}: value+grad {tg * 1e6:8.0f} us, value {tv * 1e6:7.0f} us   "
              f"K={i['patterns_per_thread']} grid={i['grid']}x{i['threads_per_cta']}", flush=True)
