"""NUTS (phylostan_b200/sampling.py): the sampler on known targets, then on the model blocks."""
import numpy as np
import pytest

from phylostan_b200 import advi, sampling

from test_advi import (CorrelatedGaussian, GaussianTarget, OracleRooted, _unconstrained_from_tree, flua_clock_problem,
                       small_problem)


def test_adaptation_windows_follow_stan():
    ends, init = sampling._windows(1000)
    assert init == 75 and ends == [100, 150, 250, 450, 950]
    ends, init = sampling._windows(150)
    assert init == 75 and ends == [100]
    ends, init = sampling._windows(100)                       # buffers do not fit: 15 % / 75 % / 10 %
    assert init == 15 and ends == [90]
    assert sampling._windows(10) == ([], 10)


def test_nuts_recovers_an_anisotropic_gaussian():
    tgt = GaussianTarget([1.0, -2.0, 0.5, 30.0], [0.5, 2.0, 1.0, 0.01])
    fit = sampling.nuts(tgt, num_warmup=600, num_samples=1500, seed=3, init="zero")
    assert fit.draws.shape == (1500, 4) and not fit.divergent.any()
    assert np.allclose(fit.draws.mean(axis=0), tgt.mean, atol=4 * tgt.sd / np.sqrt(150))
    assert np.allclose(fit.draws.std(axis=0), tgt.sd, rtol=0.15)
    assert np.allclose(np.sqrt(fit.inv_metric), tgt.sd, rtol=0.35)     # the metric learned the scales
    assert 0.6 < fit.accept_stat.mean() < 0.95 and fit.treedepth.max() <= 4
    assert fit.gradient_evaluations >= fit.n_leapfrog.sum()


def test_nuts_recovers_a_correlated_gaussian():
    cov = np.array([[1.0, 0.9], [0.9, 1.0]])
    tgt = CorrelatedGaussian([0.0, 3.0], cov)
    fit = sampling.nuts(tgt, num_warmup=500, num_samples=3000, seed=5)
    assert np.allclose(fit.draws.mean(axis=0), tgt.mean, atol=0.12)
    assert np.allclose(np.cov(fit.draws.T), cov, atol=0.15)


def test_nuts_on_a_small_unrooted_tree_agrees_with_advi():
    lik = small_problem("JC69", 1, S=5, L=60)
    m = advi.UnrootedModel(lik, "JC69")
    z0 = np.full(m.dim, -2.5)
    fit = sampling.nuts(m, num_warmup=150, num_samples=300, seed=2, init=z0)
    vb = advi.advi(m, iter=400, grad_samples=4, elbo_samples=20, eval_elbo=50, eta=0.1, seed=2, init=z0, output_samples=400)
    assert not fit.divergent.any() and fit.draws.shape == (300, lik.bcount)
    a, b = fit.draws.mean(axis=0), vb.draws.mean(axis=0)
    assert np.allclose(a, b, rtol=0.5, atol=0.02)             # same posterior, two very different approximations
    assert np.all(fit.draws > 0)


def test_nuts_runs_on_the_clock_model_block():
    d, S, lowers, heights = flua_clock_problem()
    L = 24
    lik = OracleRooted(d["peel"], d["tipmask"][:, :L].copy(), d["weights"][:L].copy(), "HKY", 4)
    m = advi.StrictClockModel(lik, "HKY", d["map"], lowers)
    fit = sampling.nuts(m, num_warmup=15, num_samples=10, seed=1, init=_unconstrained_from_tree(m, heights, lowers),
                        max_depth=4)
    assert fit.draws.shape[0] == 10 and np.all(np.isfinite(fit.lp))
    hs = fit.draws[:, [fit.names.index(f"heights.{i + 1}") for i in range(S - 1)]]
    assert np.all(hs[:, m.root - S - 1] >= hs.max(axis=1) - 1e-12)      # the root is the oldest node in every draw


@pytest.mark.gpu
def test_gpu_nuts_flua_quickstart():
    """`phylostan run ... -a nuts` on the quick-start model without Stan: the posterior means land inside the
    intervals the reference publishes for this model (README.md:103-108)."""
    from phylostan_b200 import likelihood as lk
    d, S, lowers, heights = flua_clock_problem()
    with lk.TreeLikelihood(d["peel"], d["tipmask"], d["weights"], model="HKY", categories=4, rooted=True) as lik:
        m = advi.StrictClockModel(lik, "HKY", d["map"], lowers)
        fit = sampling.nuts(m, num_warmup=300, num_samples=300, seed=4, init=_unconstrained_from_tree(m, heights, lowers),
                            max_depth=8)
    mean = fit.mean()
    assert fit.divergent.mean() < 0.05
    for name, (lo, hi) in {"wshape": (0.383, 0.616), "rate": (0.00432, 0.00577), "theta": (3.14, 5.05),
                           "kappa": (4.37, 7.039), "height": (18.36, 19.74)}.items():
        assert lo < mean[name] < hi, (name, mean[name])


def test_lockstep_hmc_recovers_a_gaussian_with_one_call_per_leapfrog_step():
    tgt = GaussianTarget([1.0, -2.0, 0.5], [0.5, 2.0, 0.05])
    calls = {"n": 0, "rows": 0}
    orig = tgt.log_prob_grad

    def counted(Z, want_grad=True):
        calls["n"] += 1
        calls["rows"] += Z.shape[0]
        return orig(Z, want_grad)

    tgt.log_prob_grad = counted
    fit = sampling.hmc(tgt, chains=16, num_warmup=300, num_samples=300, seed=2, init="zero")
    assert fit.draws.shape == (16, 300, 3) and fit.gradient_calls == calls["n"]
    assert calls["rows"] == 16 * calls["n"]                      # every call advances all chains together
    flat = fit.draws.reshape(-1, 3)
    assert np.allclose(flat.mean(axis=0), tgt.mean, atol=4 * tgt.sd / np.sqrt(400))
    assert np.allclose(flat.std(axis=0), tgt.sd, rtol=0.12)
    assert np.allclose(np.sqrt(fit.inv_metric), tgt.sd, rtol=0.3)
    assert 0.6 < fit.accept_stat.mean() <= 1.0
    # chains agree with each other (potential scale reduction close to 1)
    w = fit.draws.var(axis=1, ddof=1).mean(axis=0)
    b = fit.draws.mean(axis=1).var(axis=0, ddof=1) * 300
    assert np.all(np.sqrt(((299 / 300) * w + b / 300) / w) < 1.1), np.sqrt(((299 / 300) * w + b / 300) / w)


@pytest.mark.gpu
def test_gpu_lockstep_hmc_flua_quickstart():
    """16 chains of static HMC advanced by one phylo_b200_eval_batch per leapfrog step; the pooled posterior means
    land inside the reference's published intervals (README.md:103-108)."""
    from phylostan_b200 import likelihood as lk
    d, S, lowers, heights = flua_clock_problem()
    with lk.TreeLikelihood(d["peel"], d["tipmask"], d["weights"], model="HKY", categories=4, rooted=True) as lik:
        m = advi.StrictClockModel(lik, "HKY", d["map"], lowers)
        fit = sampling.hmc(m, chains=16, num_warmup=300, num_samples=150, seed=3,
                           init=_unconstrained_from_tree(m, heights, lowers), max_leapfrog=64)
    mean = fit.mean()
    assert fit.accept_stat.mean() > 0.5
    for name, (lo, hi) in {"wshape": (0.383, 0.616), "rate": (0.00432, 0.00577), "theta": (3.14, 5.05),
                           "kappa": (4.37, 7.039), "height": (18.36, 19.74)}.items():
        assert lo < mean[name] < hi, (name, mean[name])
