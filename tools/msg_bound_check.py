#!/usr/bin/env python3
"""How much accuracy does the message statistic cost as its bound is approached?  (GPU box only.)

Random draws on the reference's data sets whose bound (largest branch) x (largest site rate) x (eigenvalue spread) +
log(|m1|_F |m2|_F / 4) is placed at chosen values up to and beyond the library's limit of 12 (beyond it the library
falls back to the plain statistic by itself), with equal and with skewed frequencies; every gradient is compared
with the CPU oracle at the north-star tolerance 1e-8 max(1, |g|).

    python tools/msg_bound_check.py [draws per point]
"""
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)

from phylostan_b200 import encode as E, likelihood as lk  # noqa: E402
from oracle import oracle as O  # noqa: E402  (the checker)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    rng = np.random.default_rng(5)
    print("| data set | frequencies | bound | message statistic used | worst gradient error / max(1, |g|) | worst logL rel |")
    print("|---|---|---|---|---|---|")
    worst_all = 0.0
    for name, rooted in (("DS1", False), ("fluA", True), ("HCV", True)):
        z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        peel, tipmask, weights = z["peel"], z["tipmask"], z["weights"]
        S = tipmask.shape[0]
        nb = 2 * S - 2 if rooted else 2 * S - 3
        with lk.TreeLikelihood(peel, tipmask, weights, model="GTR", categories=4, rooted=rooted) as lik:
            lik.set_tiling(2, 1)   # the K = 1 kernels these small problems would get keep the plain statistic
            for skew, conc in (("equal-ish", 50.0), ("skewed", 0.4)):
                for bound in (4.0, 8.0, 10.0, 11.0, 11.9, 12.5, 20.0):
                    worst_g = worst_l = 0.0
                    used = set()
                    for _ in range(n):
                        fr = np.maximum(rng.dirichlet(np.ones(4) * conc), 1e-4)
                        fr /= fr.sum()
                        su = rng.dirichlet(np.ones(6) * 2.0)
                        rs = E.weibull_rates(rng.uniform(0.3, 1.5), 4)
                        ps = rng.dirichlet(np.ones(4) * 4)
                        dv = lk.derive("GTR", su, fr)
                        cond = max(0.0, np.log(np.linalg.norm(dv["m1"]) * np.linalg.norm(dv["m2"]) / 4.0))
                        if bound - cond <= 0.05:
                            continue
                        bl = rng.exponential(0.05, nb) + 1e-4
                        bl = np.minimum(bl, 0.5 * (bound - cond) / (np.ptp(dv["lam"]) * rs.max()))
                        bl[rng.integers(nb)] = (bound - cond) / (np.ptp(dv["lam"]) * rs.max())
                        got = lik.value_grad(bl, su, fr, rs, ps)
                        used.add(lik.info()["message_statistic"])
                        want = O.loglik_grad(peel, tipmask, weights, O.GTR, bl, su, fr, rs, ps, rooted=rooted)
                        wg = np.concatenate([want.grad_blens, want.grad_subst, want.grad_freqs, want.grad_rs, want.grad_ps])
                        worst_g = max(worst_g, float(np.max(np.abs(got.grad - wg) / np.maximum(1.0, np.abs(wg)))))
                        worst_l = max(worst_l, abs(got.log_P - want.logp) / abs(want.logp))
                    if used:
                        print(f"| {name} | {skew} | {bound} | {sorted(used)} | {worst_g:.1e} | {worst_l:.1e} |", flush=True)
                        worst_all = max(worst_all, worst_g)
    print(f"\nworst gradient error over everything: {worst_all:.1e} (tolerance 1e-8)")
    return 0 if worst_all <= 1e-8 else 1


if __name__ == "__main__":
    sys.exit(main())
