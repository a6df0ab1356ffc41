"""fluA quick start (BASELINE config 1) on the GPU without Stan: HKY + W4, heterochronous strict clock,
constant coalescent, batched mean-field ADVI.  Prints the fit summary and the time per iteration.
Usage: python tools/advi_flua.py [grad_samples] [iterations]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
from phylostan_b200 import advi, likelihood as lk  # noqa: E402
from test_advi import _unconstrained_from_tree, flua_clock_problem  # noqa: E402

gs = int(sys.argv[1]) if len(sys.argv) > 1 else 8
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
d, S, lowers, heights = flua_clock_problem()
with lk.TreeLikelihood(d["peel"], d["tipmask"], d["weights"], model="HKY", categories=4, rooted=True) as lik:
    m = advi.StrictClockModel(lik, "HKY", d["map"], lowers)
    z0 = _unconstrained_from_tree(m, heights, lowers)
    t0 = time.perf_counter()
    fit = advi.advi_meanfield(m, iter=iters, grad_samples=gs, elbo_samples=100, tol_rel_obj=0.001, seed=1, init=z0)
    dt = time.perf_counter() - t0
mean, sd = fit.mean(), dict(zip(fit.names, fit.draws.std(axis=0)))
print(f"eta {fit.eta}  iterations {fit.iterations}  converged {fit.converged}  {dt:.2f} s  "
      f"({fit.likelihood_calls} library calls, {fit.likelihood_draws} draws)")
print("ELBO", [round(e, 1) for _, e in fit.elbo_trace[:3]], "...", [round(e, 1) for _, e in fit.elbo_trace[-3:]])
for k in ("rate", "height", "theta", "kappa", "wshape", "freqs.1", "freqs.2", "freqs.3", "freqs.4"):
    print(f"  {k:8s} {mean[k]:.5g} +- {sd[k]:.3g}")
