"""Per-call latency of phylo_b200_eval_batch on DS1 (GTR+W4) as a function of the batch size.
Usage: python tools/batch_latency.py      (PHYLO_B200_NO_GRAPH=1 to compare against plain launches)"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from phylostan_b200 import encode as E, likelihood as lk  # noqa: E402

d = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "DS1.npz"))
S = d["tipmask"].shape[0]
rng = np.random.default_rng(1)
with lk.TreeLikelihood(d["peel"], d["tipmask"], d["weights"], model="GTR", categories=4, rooted=False) as lik:
    for B in (1, 8, 64, 100, 256, 8, 64):
        bl = rng.exponential(0.05, (B, 2 * S - 3)) + 1e-4
        su, fr = rng.dirichlet(np.ones(6), B), rng.dirichlet(np.ones(4) * 5, B)
        rs, ps = np.tile(E.weibull_rates(0.5, 4), (B, 1)), np.full((B, 4), 0.25)
        for grad in (True, False):
            f = lik.value_grad if grad else lik.loglik
            for _ in range(5):
                f(bl, su, fr, rs, ps)
            t0 = time.perf_counter()
            for _ in range(50):
                f(bl, su, fr, rs, ps)
            dt = (time.perf_counter() - t0) / 50
            print(f"B={B:4d} grad={int(grad)}  {dt * 1e6:8.0f} us per call   {lik.info()['patterns_per_thread']=} {lik.info()['grid']=}")
