#!/usr/bin/env python3
"""One process, N devices: the multi-device handle (phylo_b200_create_multi) on the config-3 shape, 100k patterns per
GPU (weak scaling), 64 draws per call, through the host-buffer API a Stan shim or a batched driver uses.
    python tools/multi_handle_bench.py [ndev] [B]        (GPU box with >= ndev GPUs)"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from phylostan_b200 import likelihood as lk, synth  # noqa: E402

ndev = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
S, Lper, C = 1000, 100_000, 4
base = synth.make_problem(S, 32, C, seed=synth.SEED_DATA)
rng = np.random.default_rng(synth.SEED_DATA + 5)
tm, w = synth.simulate_alignment(base.peel, base.blens, Lper * ndev, C, rng)
prob = synth.SynthProblem(S, Lper * ndev, C, base.peel, tm, w, base.blens)
draws = synth.make_draws(prob, B)
for devs in ([0], list(range(ndev))):
    L = Lper * len(devs)
    with lk.TreeLikelihood(prob.peel, tm[:, :L], w[:L], model="GTR", categories=C, devices=devs) as lik:
        for _ in range(2):
            vg = lik.value_grad(*draws)
        t0 = time.perf_counter()
        n = 3
        for _ in range(n):
            vg = lik.value_grad(*draws)
        dt = (time.perf_counter() - t0) / n
        print(f"devices {devs}: {L} patterns, {B} draws per call: {1e3 * dt:.1f} ms per call = {B / dt:.1f} tree evaluations/s, "
              f"{B * L * C / dt:.4g} pattern*category evals/s; logL[0] = {vg.log_P[0]:.6f}", flush=True)
