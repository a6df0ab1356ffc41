#!/usr/bin/env python3
"""Golden values and branch gradients from the REFERENCE's own C++ likelihood (authoring container only).

The reference's `vbsky_loglik` (eigen/eigen.j2:56-168) is rendered from the template where it lies under
/root/reference, compiled (oracle/build_ref.py; against oracle/mini_eigen because Eigen is not in the image) and run on
small problems: its 3-taxon example with every branch length 1 (eigen/eigen.cpp:5), random trees of 5-9 taxa with
ambiguous cells, JC69 with the reference's un-normalised Q (eigen/util.py:86-89), and HKY / GTR rate matrices built the
way the production Stan code builds them (generate_script.py:799-812, 855-868: Q = R diag(pi), rows summing to zero,
one expected substitution per unit time).  The template takes Q and pi as numbers, so the reference's pruning, pre-order
and gradient code runs unchanged for all of them.  Outputs: log_P and its gradient vector, whose entry i is
times[i] * dlogP/dtimes[i] (eigen.j2:165).  Writes tests/golden/ref_eigen_cpp.json.

    python tests/golden/make_golden_ref_eigen.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import build_ref  # noqa: E402

OUT = os.path.join(HERE, "ref_eigen_cpp.json")


def rate_matrix(model, subst, freqs, normalize=True):
    """generate_script.py:799-812 (HKY), :855-868 (GTR); JC69 un-normalised: eigen/util.py:86-89."""
    if model == "JC69":
        Q = np.full((4, 4), 0.25)
        np.fill_diagonal(Q, -0.75)
        return Q if not normalize else Q / 0.75
    if model == "HKY":
        k = float(subst[0])
        R = np.array([[0, 1, k, 1], [1, 0, 1, k], [k, 1, 0, 1], [1, k, 1, 0.0]])
    else:
        a, b, c, d, e, f = subst
        R = np.array([[0, a, b, c], [a, 0, d, e], [b, d, 0, f], [c, e, f, 0.0]])
    Q = R @ np.diag(freqs)
    np.fill_diagonal(Q, 0.0)
    np.fill_diagonal(Q, -Q.sum(axis=1))
    return Q / -(np.diag(Q) * freqs).sum()


def random_peel(S, rng):
    """Random rooted binary tree in this repo's encoding: rows (child, child, parent), post-order, root 2S-1."""
    live, rows, nxt = list(range(1, S + 1)), [], S + 1
    while len(live) > 1:
        i, j = sorted(rng.choice(len(live), 2, replace=False))
        a, b = live[i], live[j]
        live = [x for k, x in enumerate(live) if k not in (i, j)] + [nxt]
        rows.append((a, b, nxt))
        nxt += 1
    return np.array(rows, dtype=np.int32)


def main():
    rng = np.random.default_rng(20261019)
    cases = []

    def add(name, peel, tipmask, model, subst, freqs, normalize, blens_sets):
        Q = rate_matrix(model, subst, freqs, normalize)
        exe = build_ref.build(name, peel, tipmask, Q, freqs)
        res = build_ref.run(exe, blens_sets)
        cases.append({"name": name, "peel": peel.tolist(), "tipmask": tipmask.tolist(), "model": model,
                      "subst": [float(x) for x in subst], "freqs": [float(x) for x in freqs], "normalize": normalize,
                      "Q": Q.tolist(),
                      "runs": [{"blens": [float(x) for x in t], "log_P": lp, "grad_times_t": g.tolist()}
                               for t, (lp, g) in zip(blens_sets, res)]})

    # the reference's own example: all branch lengths 1, its un-normalised JC Q
    peel3 = np.array([[1, 2, 4], [4, 3, 5]], dtype=np.int32)
    tm3 = np.array([[1, 2, 15, 8, 4], [1, 4, 2, 8, 4], [2, 4, 1, 15, 4]], dtype=np.uint8)
    add("3tax_jc_all_ones", peel3, tm3, "JC69", [], np.full(4, 0.25), False, [[1.0] * 4, [0.1, 0.2, 0.3, 0.05]])
    for k, (S, L, model) in enumerate([(5, 12, "JC69"), (6, 20, "HKY"), (7, 24, "GTR"), (9, 40, "GTR"), (8, 30, "HKY")]):
        peel = random_peel(S, rng)
        codes = rng.choice([1, 2, 4, 8, 15], size=(S, L), p=[0.24, 0.24, 0.24, 0.24, 0.04]).astype(np.uint8)
        freqs = np.full(4, 0.25) if model == "JC69" else rng.dirichlet(np.ones(4) * 6)
        subst = [] if model == "JC69" else ([rng.lognormal(1.2, 0.4)] if model == "HKY" else rng.dirichlet(np.ones(6) * 3))
        sets = [rng.exponential(0.1, 2 * S - 2) + 1e-3, rng.exponential(0.5, 2 * S - 2) + 1e-3, np.full(2 * S - 2, 1e-4)]
        add(f"rand{k}_{model}_{S}x{L}", peel, codes, model, subst, freqs, True, sets)
    json.dump({"source": "reference eigen/eigen.j2 vbsky_loglik rendered + compiled by oracle/build_ref.py (mini_eigen)",
               "cases": cases}, open(OUT, "w"), indent=0)
    print(f"{len(cases)} cases -> {OUT}")


if __name__ == "__main__":
    main()
