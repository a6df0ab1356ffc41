#!/usr/bin/env python3
"""Golden vectors from the REFERENCE's own closed form (run in the authoring container only).

Imports /root/reference/eigen/test_ll_3tax.py unmodified.  That file needs jax, which is not
installed; ``loglik3tax`` only calls ``jnp.exp`` / ``jnp.log``, so a stub ``jax`` module whose
``numpy`` is torch lets the reference formula run in fp64 and torch autograd supplies what
``jax.value_and_grad`` (test_ll_3tax.py:329) would.  Writes tests/golden/ll_3tax.json.

    python tests/golden/make_golden_3tax.py
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference/eigen/test_ll_3tax.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ll_3tax.json")


def load_reference():
    jax = types.ModuleType("jax")
    jnp = types.ModuleType("jax.numpy")
    jnp.exp, jnp.log = torch.exp, torch.log
    jax.numpy = jnp
    sys.modules["jax"], sys.modules["jax.numpy"] = jax, jnp
    spec = importlib.util.spec_from_file_location("test_ll_3tax", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ref = load_reference()
    rng = np.random.default_rng(20261018)
    cases = []
    eye = np.eye(4)
    # (mu, pi, tips as state indices or None for ambiguous, branches b14,b24,b45,b35)
    specs = [
        (1.0, [0.25] * 4, [0, 1, 2], [1.0, 1.0, 1.0, 1.0]),     # test_ll_3tax.py:323-329, eigen.cpp:5
        (1.0, [0.25] * 4, [0, 1, 2], [0.1, 0.1, 0.2, 0.3]),     # example.tree branch lengths
        (1.0, [0.25] * 4, [2, 2, 0], [1.0, 2.0, 1.0, 1.0]),     # pruner/test.cpp column 0 shape
        (1.0, [0.25] * 4, [2, 2, 2], [1.0, 2.0, 1.0, 1.0]),     # pruner/test.cpp column 1
        (4.0 / 3.0, [0.25] * 4, [0, 3, None], [0.05, 0.4, 0.01, 0.7]),
    ]
    for _ in range(6):
        specs.append((float(rng.uniform(0.3, 2.0)), [0.25] * 4,
                      [int(x) if x < 4 else None for x in rng.integers(0, 5, 3)],
                      [float(x) for x in rng.exponential(0.3, 4) + 1e-3]))
    for mu, pi, tips, br in specs:
        tipv = np.stack([eye[t] if t is not None else np.ones(4) for t in tips])
        b = torch.tensor(br, dtype=torch.float64, requires_grad=True)
        val = ref.loglik3tax(torch.tensor(mu, dtype=torch.float64), torch.tensor(pi, dtype=torch.float64),
                             torch.tensor(tipv, dtype=torch.float64), b)
        (g,) = torch.autograd.grad(val, b)
        cases.append({"mu": mu, "pi": pi, "tips": [t if t is not None else -1 for t in tips],
                      "branches_b14_b24_b45_b35": br, "loglik": float(val), "grad": [float(x) for x in g]})
    with open(OUT, "w") as fp:
        json.dump({"source": "eigen/test_ll_3tax.py::loglik3tax via torch (see make_golden_3tax.py)",
                   "cases": cases}, fp, indent=1)
    print("wrote", OUT, len(cases), "cases; first:", cases[0]["loglik"], cases[0]["grad"])


if __name__ == "__main__":
    main()
