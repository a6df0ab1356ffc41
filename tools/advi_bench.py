"""ADVI throughput on DS1 (GTR+W4, unrooted): draws per second through the batched driver as a function
of grad_samples, i.e. what one phylo_b200_eval_batch call per iteration buys over Stan's serial loop.
Usage: python tools/advi_bench.py [iterations]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from phylostan_b200 import advi, likelihood as lk  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 300
d = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "DS1.npz"))
out = []
with lk.TreeLikelihood(d["peel"], d["tipmask"], d["weights"], model="GTR", categories=4, rooted=False) as lik:
    m = advi.UnrootedModel(lik, "GTR")
    advi.advi_meanfield(m, iter=20, grad_samples=1, elbo_samples=10, eta=0.1, seed=1, output_samples=0)   # warm-up
    for gs in (1, 8, 64, 256):
        t0 = time.perf_counter()
        fit = advi.advi_meanfield(m, iter=iters, grad_samples=gs, elbo_samples=100, eta=0.1, tol_rel_obj=1e-12,
                                  seed=1, output_samples=0)
        dt = time.perf_counter() - t0
        out.append({"grad_samples": gs, "iterations": fit.iterations, "seconds": round(dt, 3),
                    "iterations_per_s": round(fit.iterations / dt, 1),
                    "gradient_draws_per_s": round(fit.iterations * gs / dt, 1),
                    "final_elbo": round(fit.elbo_trace[-1][1], 2)})
        print(json.dumps(out[-1]), flush=True)
