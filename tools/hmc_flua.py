"""fluA quick-start model sampled by lock-step multi-chain HMC on the GPU (phylostan's `-a hmc --chains N`).
Usage: python tools/hmc_flua.py [chains] [warmup] [samples]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
from phylostan_b200 import advi, likelihood as lk, sampling  # noqa: E402
from test_advi import _unconstrained_from_tree, flua_clock_problem  # noqa: E402

chains = int(sys.argv[1]) if len(sys.argv) > 1 else 64
nw = int(sys.argv[2]) if len(sys.argv) > 2 else 500
ns = int(sys.argv[3]) if len(sys.argv) > 3 else 500
d, S, lowers, heights = flua_clock_problem()
with lk.TreeLikelihood(d["peel"], d["tipmask"], d["weights"], model="HKY", categories=4, rooted=True) as lik:
    m = advi.StrictClockModel(lik, "HKY", d["map"], lowers)
    t0 = time.perf_counter()
    fit = sampling.hmc(m, chains=chains, num_warmup=nw, num_samples=ns, seed=1,
                       init=_unconstrained_from_tree(m, heights, lowers), max_leapfrog=64)
    dt = time.perf_counter() - t0
print(f"{chains} chains x ({nw}+{ns}) iterations in {dt:.1f} s: {fit.gradient_calls} batched calls "
      f"({dt / fit.gradient_calls * 1e6:.0f} us each = {dt / fit.gradient_calls / chains * 1e6:.1f} us per chain-gradient), "
      f"step size {fit.stepsize:.4g}, up to {fit.n_leapfrog} leapfrog steps, accept {fit.accept_stat.mean():.2f}")
flat = fit.draws.reshape(-1, fit.draws.shape[-1])
for k in ("rate", "height", "theta", "kappa", "wshape"):
    col = flat[:, fit.names.index(k)]
    per = fit.draws[:, :, fit.names.index(k)]
    w, b = per.var(axis=1, ddof=1).mean(), per.mean(axis=1).var(ddof=1) * ns
    print(f"  {k:8s} mean {col.mean():.5g}  95% ({np.quantile(col, 0.025):.4g}, {np.quantile(col, 0.975):.4g})  "
          f"Rhat {np.sqrt(((ns - 1) / ns * w + b / ns) / w):.3f}")
