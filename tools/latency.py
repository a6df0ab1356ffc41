#!/usr/bin/env python3
"""Per-evaluation latency through the host-buffer C ABI on the reference's own data sets
(BASELINE configs 1, 2, 5 shapes), next to the single-thread CPU oracle.   python tools/latency.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from oracle import oracle as O  # noqa: E402
from phylostan_b200 import encode as E, likelihood as lk  # noqa: E402

G = os.path.join(os.path.dirname(__file__), "..", "tests", "golden")
for name, model, rooted in (("fluA", "HKY", True), ("DS1", "GTR", False), ("HCV", "GTR", True)):
    d = np.load(os.path.join(G, name + ".npz"))
    S, L = d["tipmask"].shape
    rng = np.random.default_rng(1)
    bl = rng.exponential(0.05, 2 * S - 2 if rooted else 2 * S - 3) + 1e-4
    subst = np.array([5.0]) if model == "HKY" else rng.dirichlet(np.ones(6))
    fr, rs, ps = rng.dirichlet(np.ones(4) * 5), E.weibull_rates(0.5, 4), np.full(4, 0.25)
    lik = lk.TreeLikelihood(d["peel"], d["tipmask"], d["weights"], model=model, categories=4, rooted=rooted)
    for _ in range(20):
        lik.value_grad(bl, subst, fr, rs, ps)
    n = 300
    t0 = time.perf_counter()
    for _ in range(n):
        lik.value_grad(bl, subst, fr, rs, ps)
    tg = (time.perf_counter() - t0) / n
    t0 = time.perf_counter()
    for _ in range(n):
        lik.loglik(bl, subst, fr, rs, ps)
    tv = (time.perf_counter() - t0) / n
    m = O.MODEL_IDS[model]
    O.loglik_grad(d["peel"], d["tipmask"], d["weights"], m, bl, subst, fr, rs, ps, rooted=rooted, nthreads=1, dp_eigen=True)
    t0 = time.perf_counter()
    for _ in range(5):
        O.loglik_grad(d["peel"], d["tipmask"], d["weights"], m, bl, subst, fr, rs, ps, rooted=rooted, nthreads=1, dp_eigen=True)
    tc = (time.perf_counter() - t0) / 5
    print(f"{name}: S={S} L={L} {model}+W4  GPU value+grad {tg * 1e6:.0f} us, value {tv * 1e6:.0f} us per call; "
          f"CPU oracle 1 thread value+grad {tc * 1e3:.2f} ms  ({tc / tg:.0f}x)  info={lik.info()['grid']}x{lik.info()['threads_per_cta']}")
    lik.set_timing(True)
    lik.value_grad(bl, subst, fr, rs, ps)
    t = lik.get_timing()
    print("   device:", {k: round(v * 1e3, 1) for k, v in t.items()}, "us")
    lik.close()
