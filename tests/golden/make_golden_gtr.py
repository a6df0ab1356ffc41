#!/usr/bin/env python3
"""Golden GTR transition matrices from the REFERENCE's own NumPy code (run in the authoring container only).

Imports /root/reference/scripts/phylo.py unmodified (Python 2 era: ``xrange`` is supplied as a builtin) and
runs its ``GTR`` class -- ``update_Q`` (:27-46), ``update_eigen_system`` (:48-61), ``p_t`` (:21-22) -- for a
dozen (rates, pi) pairs and branch lengths from 1e-9 to 50.  The rate order (a..f = AC, AG, AT, CG, CT, GT),
``Q_ij = r_ij pi_j`` and the normalisation to one substitution per unit time are the same as the
production Stan code's (phylostan/generate_script.py:855-868), so these matrices pin the GTR ``P(t)``
of the oracle and of the library against code the reference holds.  HKY is GTR with rates
(1, kappa, 1, 1, kappa, 1) (generate_script.py:799-802), so the same class pins it too.
Writes tests/golden/gtr_pt.json.

    python tests/golden/make_golden_gtr.py
"""
import builtins
import importlib.util
import json
import os

import numpy as np

REF = "/root/reference/scripts/phylo.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gtr_pt.json")


def load_reference():
    builtins.xrange = range  # scripts/phylo.py:42 is Python 2
    spec = importlib.util.spec_from_file_location("ref_phylo", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ref = load_reference()
    rng = np.random.default_rng(20261018)
    cases = []
    params = [(np.array([.10, .30, .10, .10, .30, .10]), np.array([.30, .20, .20, .30]), "survey"),
              (np.ones(6), np.full(4, 0.25), "jc-like"),                                              # degenerate spectrum
              (np.array([1, 5.58, 1, 1, 5.58, 1.0]), np.array([0.33, 0.19, 0.23, 0.25]), "hky kappa=5.58")]
    for _ in range(9):
        params.append((rng.dirichlet(np.ones(6) * 2), rng.dirichlet(np.ones(4) * 4), "random"))
    for rates, pi, tag in params:
        g = ref.GTR(list(rates), list(pi))
        g.update()
        for t in (1e-9, 1e-4, 0.02, 0.3, 1.0, 5.0, 50.0):
            P = g.p_t(t)
            cases.append({"tag": tag, "rates": list(map(float, rates)), "pi": list(map(float, pi)), "t": t,
                          "Q": np.real(g.Q).tolist(), "P": np.real(P).tolist()})
    json.dump({"source": "scripts/phylo.py GTR.update_Q/update_eigen_system/p_t (reference, executed unmodified)",
               "rate_order": "AC, AG, AT, CG, CT, GT", "cases": cases}, open(OUT, "w"), indent=0)
    print(f"wrote {len(cases)} cases to {OUT}")


if __name__ == "__main__":
    main()
