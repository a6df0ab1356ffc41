#!/usr/bin/env python3
"""Small fixed invocation for ncu: config-3 shaped problem, B draws, a few value+gradient runs and
value-only runs.   python tools/profile_run.py [B] [K] [PB] [L] [S]"""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from phylostan_b200 import likelihood as lk, synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
K = int(sys.argv[2]) if len(sys.argv) > 2 else 0
PB = int(sys.argv[3]) if len(sys.argv) > 3 else 0
L = int(sys.argv[4]) if len(sys.argv) > 4 else 100_000
S = int(sys.argv[5]) if len(sys.argv) > 5 else 1000
prob = synth.make_problem(S, L, 4, structured=False)
draws = synth.make_draws(prob, B)
lik = lk.TreeLikelihood(prob.peel, prob.tipmask, prob.weights, model="GTR", categories=4)
lik.set_tiling(K, PB)
lik.upload(*draws)
for grad in (True, True, False, False):
    lik.run(B, grad)
    lik.sync()
out = lik.download(B)
print("ok", lik.info(), out[0, 0])
