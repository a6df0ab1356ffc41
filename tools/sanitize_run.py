#!/usr/bin/env python3
"""Tiny invocation for compute-sanitizer: every kernel variant (K = 1, 2, 4; value / value+gradient;
generic and 128-thread CTAs) on a 40-taxon problem, checked against the oracle."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from oracle import oracle as O  # noqa: E402
from phylostan_b200 import encode as E, likelihood as lk, synth  # noqa: E402

for C in (4, 3):
    prob = synth.make_problem(40, 150, C, seed=3, structured=False)
    rng = np.random.default_rng(1)
    B = 2
    bl, rates, freqs, rs, ps = synth.make_draws(prob, B)
    lik = lk.TreeLikelihood(prob.peel, prob.tipmask, prob.weights, model="GTR", categories=C)
    want = [O.loglik_grad(prob.peel, prob.tipmask, prob.weights, O.GTR, bl[i], rates[i], freqs[i], rs[i], ps[i]).flat()
            for i in range(B)]
    for K, PB in ((1, 1), (2, 1), (4, 1), (1, 2)):
        lik.set_tiling(K, PB)
        vg = lik.value_grad(bl, rates, freqs, rs, ps)
        v = lik.loglik(bl, rates, freqs, rs, ps)
        for i in range(B):
            got = np.concatenate([[vg.log_P[i]], vg.grad_blens[i], vg.grad_subst[i], vg.grad_freqs[i], vg.grad_rs[i], vg.grad_ps[i]])
            assert np.max(np.abs(got - want[i]) / np.maximum(1, np.abs(want[i]))) < 1e-8
            assert abs(v[i] - want[i][0]) < 1e-9 * abs(want[i][0])
    lik.close()
print("sanitize_run ok")
