"""BASELINE config 5 without Stan: HCV, GTR + W4, heterochronous, uncorrelated-lognormal clock, skygrid
coalescent, sampled with NUTS on the GPU likelihood (ADVI first, to start the chain in the typical set).
Usage: python tools/nuts_hcv.py [warmup] [samples] [max_depth] [seed] [draws.npy]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
from phylostan_b200 import advi, likelihood as lk, sampling  # noqa: E402
from test_advi import _ucln_point, flua_clock_problem  # noqa: E402

nw = int(sys.argv[1]) if len(sys.argv) > 1 else 300
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 300
md = int(sys.argv[3]) if len(sys.argv) > 3 else 8
seed = int(sys.argv[4]) if len(sys.argv) > 4 else 1
save = sys.argv[5] if len(sys.argv) > 5 else None
d, S, lowers, heights = flua_clock_problem("HCV")
grid = np.linspace(0, 400.0, 76)[1:]                      # --grid 76 --cutoff 400
with lk.TreeLikelihood(d["peel"], d["tipmask"], d["weights"], model="GTR", categories=4, rooted=True) as lik:
    m = advi.ClockModel(lik, "GTR", d["map"], lowers, clock="ucln", coalescent="skygrid", grid=grid)
    z0 = _ucln_point(m, heights, lowers, np.random.default_rng(1))
    t0 = time.perf_counter()
    vb = advi.advi(m, iter=3000, grad_samples=8, elbo_samples=100, tol_rel_obj=0.001, seed=1, init=z0, output_samples=200)
    t1 = time.perf_counter()
    print(f"ADVI: eta {vb.eta}, {vb.iterations} iterations in {t1 - t0:.1f} s, ELBO {vb.elbo_trace[0][1]:.1f} -> "
          f"{vb.elbo_trace[-1][1]:.1f}; dim {m.dim}")
    init = vb.mu if seed == 1 else vb.draws_unconstrained[seed % len(vb.draws_unconstrained)] if hasattr(vb, "draws_unconstrained") else vb.mu
    fit = sampling.nuts(m, num_warmup=nw, num_samples=ns, seed=seed, init=init, max_depth=md)
    dt = time.perf_counter() - t1
print(f"NUTS: {nw}+{ns} iterations in {dt:.1f} s: {fit.gradient_evaluations} gradient evaluations "
      f"({dt / fit.gradient_evaluations * 1e6:.0f} us each), step size {fit.stepsize:.4g}, "
      f"mean tree depth {fit.treedepth.mean():.2f}, divergent {int(fit.divergent.sum())}")
if save:
    np.save(save, np.column_stack([fit.draws[:, fit.names.index(k)] for k in ("height", "ucln_mean", "ucln_stdev", "tau", "wshape", "thetas.1", "thetas.40", "thetas.75")]))
for k in ("height", "ucln_mean", "ucln_stdev", "tau", "wshape", "thetas.1", "thetas.40", "thetas.75"):
    col = fit.draws[:, fit.names.index(k)]
    print(f"  {k:10s} mean {col.mean():.5g}  95% ({np.quantile(col, 0.025):.4g}, {np.quantile(col, 0.975):.4g})   "
          f"[ADVI mean {vb.draws[:, vb.names.index(k)].mean():.5g}]")
