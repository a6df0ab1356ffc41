#!/usr/bin/env python3
"""Tensor-memory-stack sweep (phylo_b200_set_sweep_variant 2 / 3) against the CPU oracle and against the
shared-memory-stack kernel, then timed on the BASELINE config-3 shape (GPU box only).

    python tools/tm_check.py [B] [L] [S]      # timing shape, default 16 draws x 100 000 patterns x 1000 taxa
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)

from phylostan_b200 import likelihood as lk, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402  (the checker, not the thing measured)


def flat(vg):
    return np.concatenate([np.atleast_1d(vg.log_P).reshape(-1, 1), np.atleast_2d(vg.grad_blens), np.atleast_2d(vg.grad_subst),
                           np.atleast_2d(vg.grad_freqs), np.atleast_2d(vg.grad_rs), np.atleast_2d(vg.grad_ps)], axis=1)


def err(got, want):
    return (float(np.max(np.abs(got[:, 0] - want[:, 0]) / np.abs(want[:, 0]))),
            float(np.max(np.abs(got[:, 1:] - want[:, 1:]) / np.maximum(1.0, np.abs(want[:, 1:])))))


def parity():
    ok = True
    for name, S, L, seed, amb in (("150x700", 150, 700, 77, 0.01), ("40x130 masks", 40, 130, 5, 0.2), ("600x1500", 600, 1500, 9, 0.01)):
        prob = synth.make_problem(S, L, 4, seed=seed)
        tipmask = prob.tipmask.copy()
        if amb > 0.1:  # general masks (two-state ambiguity codes): the non-simple-tip instantiation
            rng = np.random.default_rng(3)
            sel = rng.random(tipmask.shape) < 0.1
            tipmask[sel] = np.array([3, 5, 6, 9, 10, 12, 7], dtype=np.uint8)[rng.integers(0, 7, size=int(sel.sum()))]
        B = 3
        bl, rates, freqs, rs, ps = synth.make_draws(prob, B)
        want = []
        for i in range(B):
            w = O.loglik_grad(prob.peel, tipmask, prob.weights, O.GTR, bl[i], rates[i], freqs[i], rs[i], ps[i])
            want.append(np.concatenate([[w.logp], w.grad_blens, w.grad_subst, w.grad_freqs, w.grad_rs, w.grad_ps]))
        want = np.stack(want)
        with lk.TreeLikelihood(prob.peel, tipmask, prob.weights, model="GTR", categories=4) as lik:
            lik.set_tiling(4, 1)
            for variant in (0, 3, 2):
                lik.set_sweep_variant(variant)
                got = flat(lik.value_grad(bl, rates, freqs, rs, ps))
                info = lik.info()
                e = err(got, want)
                good = e[0] <= 1e-10 and e[1] <= 1e-8 and info["sweep_variant"] == variant
                ok &= good
                print(f"[{name:13s}] variant {variant} (ran {info['sweep_variant']}, msg {info['message_statistic']} cherry {info['cherry_tables']}, slots {info['stack_slots']}/{info['stack_depth']}, "
                      f"grid {info['grid']}, smem {info['smem_bytes']}): logL rel {e[0]:.1e} grad {e[1]:.1e} {'ok' if good else 'FAIL'}",
                      flush=True)
    return ok


def timing(B, L, S):
    prob = synth.make_problem(S, L, 4, structured=False)
    draws = synth.make_draws(prob, B)
    with lk.TreeLikelihood(prob.peel, prob.tipmask, prob.weights, model="GTR", categories=4) as lik:
        lik.upload(*draws)
        base = None
        for variant in (0, 3, 2, 0, 3):
            lik.set_tiling(4, 1)
            lik.set_sweep_variant(variant)
            lik.set_timing(False)
            for _ in range(2):
                lik.run(B, True)
            lik.sync()
            lik.set_timing(True)
            ms = []
            for _ in range(3):
                lik.run(B, True)
                ms.append(lik.get_timing()["sweep_ms"])
            out = lik.download(B)
            info = lik.info()
            if base is None:
                base = out
            e = err(out, base)
            print(f"[timing {S}x{L}x{B}] variant {variant} (ran {info['sweep_variant']}, msg {info['message_statistic']} cherry {info['cherry_tables']}, slots {info['stack_slots']}/{info['stack_depth']}, "
                  f"grid {info['grid']}): sweep {min(ms):.2f} ms -> {B / min(ms) * 1e3:.1f} evals/s; vs first: logL rel {e[0]:.1e} "
                  f"grad {e[1]:.1e}", flush=True)


if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
    S = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
    t0 = time.time()
    ok = parity()
    print(f"parity {'ok' if ok else 'FAILED'} ({time.time() - t0:.0f} s)", flush=True)
    if ok or os.environ.get("TM_TIME_ANYWAY"):
        timing(B, L, S)
    sys.exit(0 if ok else 1)
