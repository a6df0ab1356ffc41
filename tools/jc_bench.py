"""JC69 value+gradient throughput on the config-3 shape (1000 taxa x 100k patterns) and latency on DS1:
scalar-statistic sweep (default) vs the generic 4x4 statistics (PHYLO_B200_NO_JC_SCALAR=1).
Usage: python tools/jc_bench.py [B]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from phylostan_b200 import encode as E, likelihood as lk, synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
prob = synth.make_problem(1000, 100_000, 4, structured=False)
bl, rates, freqs, rs, ps = synth.make_draws(prob, B)
for C in (4, 1):
    with lk.TreeLikelihood(prob.peel, prob.tipmask, prob.weights, model="JC69", categories=C) as lik:
        r, p = (rs, ps) if C == 4 else (np.ones((B, 1)), np.ones((B, 1)))
        lik.upload(bl, None, None, r, p)
        for _ in range(2):
            lik.run(B, True)
        lik.sync()
        lik.set_timing(True)
        ms = []
        for _ in range(3):
            lik.run(B, True)
            ms.append(lik.get_timing()["sweep_ms"])
        print(f"JC69 C={C} 1000x100k B={B}: sweep {min(ms):.2f} ms -> {B / min(ms) * 1e3:.1f} evals/s  K={lik.info()['patterns_per_thread']}", flush=True)
d = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "DS1.npz"))
S = d["tipmask"].shape[0]
rng = np.random.default_rng(1)
b1 = rng.exponential(0.05, 2 * S - 3) + 1e-4
with lk.TreeLikelihood(d["peel"], d["tipmask"], d["weights"], model="JC69", categories=1, rooted=False) as lik:
    for _ in range(10):
        lik.value_grad(b1, None, None, np.ones(1), np.ones(1))
    t0 = time.perf_counter()
    for _ in range(200):
        lik.value_grad(b1, None, None, np.ones(1), np.ones(1))
    print(f"DS1 JC69 C=1 single evaluation: {(time.perf_counter() - t0) / 200 * 1e6:.0f} us")
