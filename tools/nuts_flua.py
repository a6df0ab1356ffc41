"""fluA quick-start model sampled with NUTS on the GPU (the reference's `-a nuts` run).
Usage: python tools/nuts_flua.py [warmup] [samples]"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
from phylostan_b200 import advi, likelihood as lk, sampling  # noqa: E402
from test_advi import _unconstrained_from_tree, flua_clock_problem  # noqa: E402

nw = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
d, S, lowers, heights = flua_clock_problem()
with lk.TreeLikelihood(d["peel"], d["tipmask"], d["weights"], model="HKY", categories=4, rooted=True) as lik:
    m = advi.StrictClockModel(lik, "HKY", d["map"], lowers)
    t0 = time.perf_counter()
    fit = sampling.nuts(m, num_warmup=nw, num_samples=ns, seed=1, init=_unconstrained_from_tree(m, heights, lowers))
    dt = time.perf_counter() - t0
print(f"{nw}+{ns} iterations in {dt:.1f} s: {fit.gradient_evaluations} gradient evaluations "
      f"({dt / fit.gradient_evaluations * 1e6:.0f} us each), step size {fit.stepsize:.4g}, "
      f"mean tree depth {fit.treedepth.mean():.2f}, divergent {int(fit.divergent.sum())}")
mean = fit.mean()
import numpy as np  # noqa: E402
for k in ("rate", "height", "theta", "kappa", "wshape"):
    col = fit.draws[:, fit.names.index(k)]
    print(f"  {k:8s} mean {mean[k]:.5g}  95% ({np.quantile(col, 0.025):.4g}, {np.quantile(col, 0.975):.4g})")
