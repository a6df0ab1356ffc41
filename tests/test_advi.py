"""Batched ADVI driver (phylostan_b200/advi.py): transforms, the model block's gradient, the optimiser.
CPU tests drive the model block with the oracle as likelihood back end; the GPU tests with the library."""
import math

import numpy as np
import pytest

from oracle import oracle as O
from phylostan_b200 import advi, synth
from phylostan_b200 import encode as E

from conftest import GOLDEN


class OracleLikelihood:
    """The TreeLikelihood surface UnrootedModel uses, answered by the CPU oracle (test infrastructure)."""

    def __init__(self, peel, tipmask, weights, model, C):
        self.peel, self.tipmask, self.weights = peel, tipmask, weights
        self.model, self.C = {"JC69": O.JC69, "HKY": O.HKY, "GTR": O.GTR}[model], C
        self.bcount = 2 * tipmask.shape[0] - 3
        self.nsubst = {"JC69": 0, "HKY": 1, "GTR": 6}[model]
        self.calls = 0

    def _each(self, blens, subst, freqs, rs, ps, want_grad):
        self.calls += 1
        out = []
        for b in range(blens.shape[0]):
            out.append(O.loglik_grad(self.peel, self.tipmask, self.weights, self.model, blens[b],
                                     None if subst is None else subst[b], None if freqs is None else freqs[b],
                                     rs[b], ps[b], rooted=False, want_grad=want_grad))
        return out

    def value_grad(self, blens, subst=None, freqs=None, rs=None, ps=None):
        from phylostan_b200.likelihood import ValueGrad
        r = self._each(blens, subst, freqs, rs, ps, True)
        st = lambda name: np.stack([np.atleast_1d(getattr(x, name)) for x in r])
        return ValueGrad(np.array([x.logp for x in r]), st("grad_blens"),
                         st("grad_subst") if self.nsubst else np.zeros((len(r), 0)), st("grad_freqs"), st("grad_rs"),
                         st("grad_ps"))

    def loglik(self, blens, subst=None, freqs=None, rs=None, ps=None):
        return np.array([x.logp for x in self._each(blens, subst, freqs, rs, ps, False)])


def small_problem(model, C, S=6, L=40, seed=5):
    prob = synth.make_problem(S, L, C, seed=seed)
    peel = E.unrooted_swap(prob.peel)
    return OracleLikelihood(peel, prob.tipmask, prob.weights, model, C)


def test_simplex_transform_is_stans_stick_breaking():
    rng = np.random.default_rng(0)
    y = rng.normal(0, 1.5, (7, 5))
    x, logj, cache = advi.simplex_constrain(y)
    assert np.allclose(x.sum(axis=1), 1.0) and np.all(x > 0)
    assert np.allclose(advi.simplex_constrain(np.zeros((1, 3)))[0], 0.25)       # y = 0 is the uniform simplex
    # log|J| against the determinant of the numerical Jacobian of the first K-1 coordinates
    for b in range(3):
        J = np.zeros((5, 5))
        for k in range(5):
            e = np.zeros(5); e[k] = 1e-6
            J[:, k] = (advi.simplex_constrain((y[b] + e)[None])[0][0, :5] - advi.simplex_constrain((y[b] - e)[None])[0][0, :5]) / 2e-6
        assert math.log(abs(np.linalg.det(J))) == pytest.approx(logj[b], abs=1e-6)
    # adjoint: d/dy [ sum(c * log x) + logJ ]
    c = rng.uniform(0.5, 2.0, 6)
    f = lambda yy: (c * np.log(advi.simplex_constrain(yy)[0])).sum(axis=1) + advi.simplex_constrain(yy)[1]
    g = advi.simplex_adjoint(c / x, cache)
    for k in range(5):
        e = np.zeros(5); e[k] = 1e-6
        assert np.allclose((f(y + e) - f(y - e)) / 2e-6, g[:, k], rtol=1e-6, atol=1e-7)


def test_weibull_rates_and_derivative():
    w = np.array([0.488, 1.3])
    rs, drs = advi.weibull_rates(w, 4)
    assert np.allclose(rs[0], E.weibull_rates(0.488, 4)) and np.allclose(rs.mean(axis=1), 1.0)
    fd = (advi.weibull_rates(w + 1e-6, 4)[0] - advi.weibull_rates(w - 1e-6, 4)[0]) / 2e-6
    assert np.allclose(drs, fd, rtol=1e-6, atol=1e-8)


@pytest.mark.parametrize("model,C", [("GTR", 4), ("HKY", 4), ("JC69", 1), ("JC69", 3)])
def test_model_block_gradient_matches_finite_differences(model, C):
    lik = small_problem(model, C)
    m = advi.UnrootedModel(lik, model, rates_alpha=np.array([1, 2, 1, 1, 2, 1.5]), freqs_alpha=np.array([2, 1, 1, 3.0]))
    assert m.dim == (1 if C > 1 else 0) + lik.bcount + {"GTR": 8, "HKY": 4, "JC69": 0}[model]
    assert len(m.constrained_names()) == m.constrained_matrix(np.zeros((1, m.dim))).shape[1]
    rng = np.random.default_rng(3)
    Z = rng.normal(-1.0, 0.7, (3, m.dim))
    lp, G = m.log_prob_grad(Z)
    assert lik.calls == 1                                      # one batched likelihood call for all draws
    assert np.allclose(lp, m.log_prob(Z), rtol=1e-13)
    for k in range(m.dim):
        e = np.zeros(m.dim); e[k] = 1e-6
        fd = (m.log_prob(Z + e) - m.log_prob(Z - e)) / 2e-6
        assert np.allclose(fd, G[:, k], rtol=2e-5, atol=2e-5), (k, fd, G[:, k])


def test_model_block_value_is_the_stan_program():
    """log_prob restated literally from tests/golden/DS1-GTR-W4-external.stan for one draw."""
    lik = small_problem("GTR", 4)
    m = advi.UnrootedModel(lik, "GTR")
    rng = np.random.default_rng(11)
    z = rng.normal(-1.0, 0.5, m.dim)
    c = m.constrain(z[None])
    wshape, blens, rates, freqs = c["wshape"][0], c["blens"][0], c["rates"][0], c["freqs"][0]
    assert wshape == pytest.approx(0.1 + math.exp(z[0]))
    C = 4
    rs = np.array([(-math.log(1.0 - (2.0 * i + 1.0) / (2.0 * C))) ** (1.0 / wshape) for i in range(C)])
    rs /= rs.sum() / C
    target = -wshape - 10.0 * blens.sum()                    # exponential(1), exponential(10); dirichlet(1) is flat
    target += O.loglik_grad(lik.peel, lik.tipmask, lik.weights, O.GTR, blens, rates, freqs, rs, np.full(C, 1.0 / C),
                            rooted=False, want_grad=False).logp
    assert m.log_prob(z[None])[0] - c["logj"][0] == pytest.approx(target, rel=1e-12)
    # draws outside the domain are dropped, not fatal
    bad = np.vstack([z, np.full(m.dim, 800.0)])
    lp = m.log_prob(bad)
    assert np.isfinite(lp[0]) and lp[1] == -np.inf


class GaussianTarget:
    """log p(z) = -1/2 sum ((z - m)/s)^2: the mean-field optimum is mu = m, omega = log s."""

    def __init__(self, mean, sd):
        self.mean, self.sd, self.dim = np.asarray(mean, float), np.asarray(sd, float), len(mean)

    def log_prob(self, Z):
        return -0.5 * (((Z - self.mean) / self.sd) ** 2).sum(axis=1)

    def log_prob_grad(self, Z, want_grad=True):
        return self.log_prob(Z), -(Z - self.mean) / self.sd ** 2

    def constrained_names(self):
        return [f"z.{i + 1}" for i in range(self.dim)]

    def constrained_matrix(self, Z):
        return Z


def test_advi_recovers_a_gaussian():
    tgt = GaussianTarget([1.0, -2.0, 0.5, 3.0], [0.5, 2.0, 1.0, 0.1])
    fit = advi.advi_meanfield(tgt, iter=3000, grad_samples=16, elbo_samples=200, tol_rel_obj=1e-4, seed=4, init="zero")
    assert fit.eta in advi._Advi.ETA_SEQUENCE
    assert np.allclose(fit.mu, tgt.mean, atol=0.15)
    assert np.allclose(np.exp(fit.omega), tgt.sd, rtol=0.2)
    assert fit.elbo_trace[-1][1] > fit.elbo_trace[0][1]
    assert fit.draws.shape == (1000, 4)
    # one library call per iteration / per ELBO estimate, regardless of the number of draws
    assert fit.likelihood_draws > 10 * fit.likelihood_calls


def test_advi_on_a_small_tree_with_the_oracle_back_end():
    lik = small_problem("JC69", 1, S=5, L=30)
    m = advi.UnrootedModel(lik, "JC69")
    fit = advi.advi_meanfield(m, iter=60, grad_samples=2, elbo_samples=20, eval_elbo=20, eta=0.1, seed=2,
                              init=np.full(m.dim, -2.0), output_samples=50)
    assert fit.iterations <= 60 and len(fit.elbo_trace) >= 2
    assert fit.elbo_trace[-1][1] > fit.elbo_trace[0][1]
    assert fit.draws.shape == (50, lik.bcount) and np.all(fit.draws > 0)


# --------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_gpu_model_block_matches_oracle_back_end_on_ds1():
    from phylostan_b200 import likelihood as lk
    d = np.load(GOLDEN + "/DS1.npz")
    rng = np.random.default_rng(8)
    ora = OracleLikelihood(d["peel"], d["tipmask"], d["weights"], "GTR", 4)
    with lk.TreeLikelihood(d["peel"], d["tipmask"], d["weights"], model="GTR", categories=4, rooted=False) as lik:
        m_gpu, m_cpu = advi.UnrootedModel(lik, "GTR"), advi.UnrootedModel(ora, "GTR")
        Z = rng.normal(-2.5, 0.5, (6, m_gpu.dim))
        lp, G = m_gpu.log_prob_grad(Z)
        lp0, G0 = m_cpu.log_prob_grad(Z)
        assert np.max(np.abs(lp - lp0) / np.abs(lp0)) <= 1e-10
        assert np.max(np.abs(G - G0) / np.maximum(1.0, np.abs(G0))) <= 1e-8
        assert np.max(np.abs(m_gpu.log_prob(Z) - lp0) / np.abs(lp0)) <= 1e-10


@pytest.mark.gpu
def test_gpu_batched_advi_on_ds1():
    """DS1, GTR+W4 (BASELINE config 2's model): the ELBO rises and the fit lands on sensible values."""
    from phylostan_b200 import likelihood as lk
    d = np.load(GOLDEN + "/DS1.npz")
    with lk.TreeLikelihood(d["peel"], d["tipmask"], d["weights"], model="GTR", categories=4, rooted=False) as lik:
        m = advi.UnrootedModel(lik, "GTR")
        fit = advi.advi_meanfield(m, iter=1500, grad_samples=8, elbo_samples=100, tol_rel_obj=0.001, seed=3,
                                  output_samples=200)
    e0, e1 = fit.elbo_trace[0][1], fit.elbo_trace[-1][1]
    assert e1 > e0 and e1 > -9000.0            # the alignment's log-likelihood at a fitted tree is about -7100
    mean = fit.mean()
    total = sum(v for k, v in mean.items() if k.startswith("blens."))
    assert 0.5 < total < 10.0
    assert abs(sum(mean[f"freqs.{i}"] for i in range(1, 5)) - 1.0) < 1e-9
    assert mean["rates.2"] > mean["rates.1"] and mean["rates.5"] > mean["rates.6"]    # transitions > transversions
    assert fit.likelihood_calls < fit.iterations + fit.iterations // 100 + 400


# --------------------------------------------------------------------------------------------------- clock trees
def flua_clock_problem(name="fluA"):
    """Golden data + tip dates / lower bounds derived from the time tree (utils.py:5-16, 93-104)."""
    d = np.load(GOLDEN + "/" + name + ".npz")
    S = d["tipmask"].shape[0]
    nn = 2 * S - 1
    depth = {nn: 0.0}
    for node, par in d["map"][1:]:
        depth[int(node)] = depth[int(par)] + float(d["tree_blens"][int(node) - 1])
    top = max(depth[k] for k in range(1, S + 1))
    lowers = np.zeros(nn)
    for k in range(1, S + 1):
        lowers[k - 1] = top - depth[k]
    for node, par in d["map"][:0:-1]:                         # reverse pre-order: children before parents
        lowers[int(par) - 1] = max(lowers[int(par) - 1], lowers[int(node) - 1])
    heights = np.array([top - depth[S + 1 + k] for k in range(S - 1)])
    return d, S, lowers, heights


class OracleRooted(OracleLikelihood):
    def __init__(self, peel, tipmask, weights, model, C):
        super().__init__(peel, tipmask, weights, model, C)
        self.bcount = 2 * tipmask.shape[0] - 2

    def _each(self, blens, subst, freqs, rs, ps, want_grad):
        self.calls += 1
        return [O.loglik_grad(self.peel, self.tipmask, self.weights, self.model, blens[b],
                              None if subst is None else subst[b], None if freqs is None else freqs[b], rs[b], ps[b],
                              rooted=True, want_grad=want_grad) for b in range(blens.shape[0])]


def _unconstrained_from_tree(m, heights, lowers, rate=0.004, theta=5.0, wshape=0.6, kappa=4.0):
    """Unconstrained point whose heights are the time tree's (inverse of the ratio transform)."""
    z = np.zeros(m.dim)
    h = heights
    props = np.array([(h[m.tr_node[j]] - m.tr_lo[j]) / (h[m.tr_parent[j]] - m.tr_lo[j]) for j in range(m.S - 2)])
    props = np.clip(props, 1e-6, 1 - 1e-6)
    z[m.slices["props"]] = np.log(props) - np.log1p(-props)
    z[m.slices["rate"]] = math.log(rate)
    z[m.slices["height"]] = math.log(h[m.root - m.S - 1] - m.lower_root)
    z[m.slices["theta"]] = math.log(theta)
    if "wshape" in m.slices:
        z[m.slices["wshape"]] = math.log(wshape - 0.1)
    if "kappa" in m.slices:
        z[m.slices["kappa"]] = math.log(kappa)
    return z


def test_strict_clock_model_is_the_stan_program():
    """One draw of tests/golden/fluA-HKY-W4-external.stan restated literally: transform(), the
    heights -> blens loop, constant_coalescent_log, the priors and the log-det-Jacobian loop."""
    d, S, lowers, heights = flua_clock_problem()
    lik = OracleRooted(d["peel"], d["tipmask"], d["weights"], "HKY", 4)
    m = advi.StrictClockModel(lik, "HKY", d["map"], lowers)
    assert m.dim == 1 + (S - 2) + 3 + 1 + 3
    z = _unconstrained_from_tree(m, heights, lowers) + np.random.default_rng(0).normal(0, 0.05, m.dim)
    c = m.constrain(z[None])
    props, rate, height, theta = c["props"][0], c["rate"][0], c["height"][0], c["theta"][0]
    kappa, freqs, wshape = c["kappa"][0], c["freqs"][0], c["wshape"][0]
    mp = [[int(a), int(b)] for a, b in d["map"]]
    nodeCount = 2 * S - 1
    # transform (generate_script.py:711-735)
    hs = [0.0] * (S - 1)
    hs[mp[0][0] - S - 1] = height
    j = 0
    for i in range(1, nodeCount):
        if mp[i][0] > S:
            lo = lowers[mp[i][0] - 1]
            hs[mp[i][0] - S - 1] = lo + (hs[mp[i][1] - S - 1] - lo) * props[j]
            j += 1
    assert np.allclose(hs, c["heights"][0], rtol=1e-14)
    # heights -> blens (generate_script.py:660-679)
    blens = np.zeros(2 * S - 2)
    for i in range(1, nodeCount):
        node, par = mp[i]
        blens[node - 1] = rate * (hs[par - S - 1] - (hs[node - S - 1] if node > S else lowers[node - 1]))
    # constant_coalescent_log (generate_script.py:285-349)
    times = [lowers[k] for k in range(S)] + hs
    child = [0] * S + [2] * (S - 1)
    order = sorted(range(nodeCount), key=lambda k: times[k])
    logP, lineages, start = 0.0, 0.0, times[order[0]]
    for k in order:
        interval = times[k] - start
        if interval != 0.0:
            logP -= interval * (lineages * (lineages - 1.0)) / 2.0 / theta
        if child[k] == 0:
            lineages += 1.0
        else:
            lineages -= 1.0
            logP -= math.log(theta)
        start = times[k]
    C = 4
    rs = np.array([(-math.log(1.0 - (2.0 * i + 1.0) / (2.0 * C))) ** (1.0 / wshape) for i in range(C)])
    rs /= rs.sum() / C
    target = -wshape - 1000.0 * rate - math.log(theta) + logP
    lk = math.log(kappa)
    target += -lk - (lk - 1.0) ** 2 / (2 * 1.25 ** 2)
    target += O.loglik_grad(d["peel"], d["tipmask"], d["weights"], O.HKY, blens, np.array([kappa]), freqs, rs,
                            np.full(C, 0.25), rooted=True, want_grad=False).logp
    for i in range(1, nodeCount):
        if mp[i][0] > S:
            target += math.log(hs[mp[i][1] - S - 1] - lowers[mp[i][0] - 1])
    assert m.log_prob(z[None])[0] - c["logj"][0] == pytest.approx(target, rel=1e-12)


def test_strict_clock_model_gradient_matches_finite_differences():
    d, S, lowers, heights = flua_clock_problem()
    L = 40                                                    # a slice of the patterns keeps the oracle fast
    lik = OracleRooted(d["peel"], d["tipmask"][:, :L].copy(), d["weights"][:L].copy(), "HKY", 4)
    m = advi.StrictClockModel(lik, "HKY", d["map"], lowers, freqs_alpha=np.array([2.0, 1, 1, 3]))
    rng = np.random.default_rng(4)
    Z = _unconstrained_from_tree(m, heights, lowers)[None] + rng.normal(0, 0.1, (2, m.dim))
    lp, G = m.log_prob_grad(Z)
    assert np.all(np.isfinite(lp)) and lik.calls == 1
    for k in range(m.dim):
        e = np.zeros(m.dim); e[k] = 1e-6
        fd = (m.log_prob(Z + e) - m.log_prob(Z - e)) / 2e-6
        assert np.allclose(fd, G[:, k], rtol=5e-5, atol=5e-5), (k, fd, G[:, k])
    # contemporaneous tips, JC69, one category: the other branch of every `if heterochronous`
    lik1 = OracleRooted(d["peel"], d["tipmask"][:, :L].copy(), d["weights"][:L].copy(), "JC69", 1)
    m1 = advi.StrictClockModel(lik1, "JC69", d["map"])
    assert m1.lower_root == 0.0 and m1.dim == (S - 2) + 3
    Z1 = rng.normal(0, 0.3, (2, m1.dim))
    Z1[:, m1.slices["rate"]] -= 5.0
    lp1, G1 = m1.log_prob_grad(Z1)
    for k in range(0, m1.dim, 3):
        e = np.zeros(m1.dim); e[k] = 1e-6
        fd = (m1.log_prob(Z1 + e) - m1.log_prob(Z1 - e)) / 2e-6
        assert np.allclose(fd, G1[:, k], rtol=5e-5, atol=5e-5), (k, fd, G1[:, k])


@pytest.mark.gpu
def test_gpu_flua_quickstart_batched_advi():
    """BASELINE config 1 without Stan: fluA, HKY + W4, heterochronous strict clock, constant coalescent,
    mean-field ADVI with all draws of an iteration in one library call."""
    from phylostan_b200 import likelihood as lk
    d, S, lowers, heights = flua_clock_problem()
    ora = OracleRooted(d["peel"], d["tipmask"], d["weights"], "HKY", 4)
    with lk.TreeLikelihood(d["peel"], d["tipmask"], d["weights"], model="HKY", categories=4, rooted=True) as lik:
        m = advi.StrictClockModel(lik, "HKY", d["map"], lowers)
        z0 = _unconstrained_from_tree(m, heights, lowers)
        Z = z0[None] + np.random.default_rng(1).normal(0, 0.1, (4, m.dim))
        lp, G = m.log_prob_grad(Z)
        lp0, G0 = advi.StrictClockModel(ora, "HKY", d["map"], lowers).log_prob_grad(Z)
        assert np.max(np.abs(lp - lp0) / np.abs(lp0)) <= 1e-10
        assert np.max(np.abs(G - G0) / np.maximum(1.0, np.abs(G0))) <= 1e-7
        fit = advi.advi_meanfield(m, iter=10000, grad_samples=8, elbo_samples=100, tol_rel_obj=0.001, seed=5, init=z0,
                                  output_samples=1000)
    mean = fit.mean()
    assert fit.converged and fit.elbo_trace[-1][1] > fit.elbo_trace[0][1]
    # The reference publishes this run's outcome (README.md:103-108, `phylostan run ... -q meanfield`):
    # posterior means with 95 % intervals.  Every mean of the batched driver must fall inside them.
    published = {"wshape": (0.488, 0.383, 0.616), "rate": (0.00499, 0.00432, 0.00577), "theta": (4.03, 3.14, 5.05),
                 "kappa": (5.58, 4.37, 7.039), "height": (18.96, 18.36, 19.74)}
    for name, (ref_mean, lo, hi) in published.items():
        assert lo < mean[name] < hi, (name, mean[name], ref_mean)
        assert abs(mean[name] - ref_mean) < 0.5 * (hi - lo), (name, mean[name], ref_mean)
    assert abs(sum(mean[f"freqs.{i}"] for i in range(1, 5)) - 1.0) < 1e-9


def test_setup_dates_and_lowers_from_the_time_tree():
    """encode.setup_dates / get_lowers (utils.py:5-57, 93-104) on the reference's fluA files."""
    import os
    ref = "/root/reference/examples/fluA"
    if not os.path.exists(ref):
        pytest.skip("reference tree not mounted (GPU box)")
    tree = E.read_tree(ref + "/fluA.tree")
    enc = E.encode(tree, E.read_alignment(ref + "/fluA.fa"), rooted=True)
    oldest = E.setup_dates(tree, None, True)
    lowers = E.get_lowers(tree)
    d, S, want, heights = flua_clock_problem()
    assert np.array_equal(enc.map, d["map"])
    assert np.allclose(lowers, want, atol=1e-9) and oldest == pytest.approx(want.max(), abs=1e-9)
    assert lowers[:S].min() == pytest.approx(0.0, abs=1e-9)          # the most recent tip defines time 0
    # contemporaneous: all dates zero, no lower bounds
    assert E.setup_dates(tree, None, False) is None and E.get_lowers(tree).max() == 0.0
    # dates given explicitly as calendar years
    years = {n.label: 2000.0 + i % 7 for i, n in enumerate(tree.leaves())}
    assert E.setup_dates(tree, years, False) == pytest.approx(6.0)
    assert max(n.date for n in tree.leaves()) == pytest.approx(6.0) and min(n.date for n in tree.leaves()) == 0.0


class CorrelatedGaussian(GaussianTarget):
    """log p(z) = -1/2 (z - m)^T P (z - m): the full-rank optimum is mu = m, L L^T = P^-1."""

    def __init__(self, mean, cov):
        self.mean, self.cov, self.dim = np.asarray(mean, float), np.asarray(cov, float), len(mean)
        self.prec = np.linalg.inv(self.cov)

    def log_prob(self, Z):
        dz = Z - self.mean
        return -0.5 * np.einsum("bi,ij,bj->b", dz, self.prec, dz)

    def log_prob_grad(self, Z, want_grad=True):
        return self.log_prob(Z), -(Z - self.mean) @ self.prec


def test_fullrank_advi_recovers_a_correlated_gaussian():
    cov = np.array([[1.0, 0.8, 0.0], [0.8, 1.0, -0.3], [0.0, -0.3, 0.5]])
    tgt = CorrelatedGaussian([0.5, -1.0, 2.0], cov)
    fit = advi.advi(tgt, algorithm="fullrank", iter=4000, grad_samples=32, elbo_samples=200, tol_rel_obj=1e-5, seed=7,
                    init="zero")
    assert fit.omega is None and fit.L.shape == (3, 3) and np.allclose(np.triu(fit.L, 1), 0.0)
    assert np.allclose(fit.mu, tgt.mean, atol=0.15)
    assert np.allclose(fit.L @ fit.L.T, cov, atol=0.2)
    # the mean-field family cannot represent the correlation: its variances are the conditional ones, 1 / P_ii
    mf = advi.advi(tgt, algorithm="meanfield", iter=4000, grad_samples=32, elbo_samples=200, tol_rel_obj=1e-5, seed=7,
                   init="zero")
    assert np.allclose(np.exp(2 * mf.omega), 1.0 / np.diag(tgt.prec), rtol=0.3)
    assert fit.elbo_trace[-1][1] > mf.elbo_trace[-1][1] - 0.05
    with pytest.raises(ValueError):
        advi.advi(tgt, algorithm="lowrank")


# --------------------------------------------------------------------------------------------------- ucln + skygrid
def _hcv_model(lik_cls=None, L=None, model="GTR", C=4):
    d, S, lowers, heights = flua_clock_problem("HCV")
    tm, w = (d["tipmask"], d["weights"]) if L is None else (d["tipmask"][:, :L].copy(), d["weights"][:L].copy())
    lik = OracleRooted(d["peel"], tm, w, model, C)
    grid = np.linspace(0, 400.0, 6)[1:]                        # phylostan.py:275-277 with --grid 6 --cutoff 400
    m = advi.ClockModel(lik, model, d["map"], lowers, clock="ucln", coalescent="skygrid", grid=grid)
    return d, S, lowers, heights, lik, grid, m


def _ucln_point(m, heights, lowers, rng):
    z = rng.normal(0, 0.1, m.dim)
    h = heights
    props = np.array([(h[m.tr_node[j]] - m.tr_lo[j]) / (h[m.tr_parent[j]] - m.tr_lo[j]) for j in range(m.S - 2)])
    props = np.clip(props, 1e-4, 1 - 1e-4)
    z[m.slices["props"]] += np.log(props) - np.log1p(-props)
    z[m.slices["height"]] += math.log(h[m.root - m.S - 1] - m.lower_root)
    z[m.slices["substrates"]] += math.log(8e-4)
    z[m.slices["ucln_mean"]] += math.log(8e-4)
    z[m.slices["ucln_stdev"]] += math.log(0.4)
    z[m.slices["thetas"]] += 5.0
    z[m.slices["wshape"]] += math.log(0.4)
    return z


def test_ucln_skygrid_model_is_the_stan_program():
    """BASELINE config 5's program (ucln clock + skygrid, HCV) for one draw, restated literally from the
    generator's text: skygrid_coalescent_log, gmrf_log, the lognormal clock prior, the per-branch rates."""
    d, S, lowers, heights, lik, grid, m = _hcv_model(L=30)
    G = grid.size
    assert m.dim == 1 + (S - 2) + (2 * S - 2) + 2 + 1 + G + 1 + 5 + 3
    z = _ucln_point(m, heights, lowers, np.random.default_rng(3))
    c = m.constrain(z[None])
    hs, sub, mean, sd = list(c["heights"][0]), c["substrates"][0], c["ucln_mean"][0], c["ucln_stdev"][0]
    pop, tau, wshape, rates, freqs = c["thetas"][0], c["tau"][0], c["wshape"][0], c["rates"][0], c["freqs"][0]
    mp = [[int(a), int(b)] for a, b in d["map"]]
    nodeCount = 2 * S - 1
    blens = np.zeros(2 * S - 2)
    for i in range(1, nodeCount):
        node, par = mp[i]
        blens[node - 1] = sub[node - 1] * (hs[par - S - 1] - (hs[node - S - 1] if node > S else lowers[node - 1]))
    # skygrid_coalescent_log, 1-based indices of the Stan text mapped to 0-based
    times = [lowers[k] for k in range(S)] + hs
    child = [0] * S + [2] * (S - 1)
    order = sorted(range(nodeCount), key=lambda k: times[k])
    logP, index, lineages, start = 0.0, 1, 0.0, times[order[0]]
    logPop = pop[index - 1]; popSize = math.exp(logPop)
    for k in order:
        finish = times[k]
        while index < G and finish > grid[index - 1]:
            end = min(grid[index - 1], finish)
            logP -= (end - start) * (lineages * (lineages - 1.0)) / 2.0 / popSize
            start = end
            if index < G:
                index += 1
                logPop = pop[index - 1]; popSize = math.exp(logPop)
        logP -= (finish - start) * (lineages * (lineages - 1.0)) / 2.0 / popSize
        if child[k] != 0:
            logP -= logPop
        lineages += 1.0 if child[k] == 0 else -1.0
        start = finish
    gmrf = math.log(tau) * (G - 1.0) / 2.0 - sum((pop[i] - pop[i - 1]) ** 2 for i in range(1, G)) * tau / 2.0
    C = 4
    rs = np.array([(-math.log(1.0 - (2.0 * i + 1.0) / (2.0 * C))) ** (1.0 / wshape) for i in range(C)])
    rs /= rs.sum() / C
    mu = math.log(mean) - sd * sd * 0.5
    target = -wshape
    target += sum(-math.log(sd) - math.log(x) - (math.log(x) - mu) ** 2 / (2 * sd * sd) for x in sub)   # lognormal
    target += -1000.0 * mean + (0.5396 - 1.0) * math.log(sd) - 2.6184 * sd
    target += logP + gmrf + (0.001 - 1.0) * math.log(tau) - 0.001 * tau
    target += O.loglik_grad(lik.peel, lik.tipmask, lik.weights, O.GTR, blens, rates, freqs, rs, np.full(C, 0.25),
                            rooted=True, want_grad=False).logp
    for i in range(1, nodeCount):
        if mp[i][0] > S:
            target += math.log(hs[mp[i][1] - S - 1] - lowers[mp[i][0] - 1])
    # gmrf_log keeps its -(G-1)/2 log(2 pi) constant (a user-defined function, not a built-in `~`)
    target += -(G - 1.0) / 2.0 * math.log(2.0 * math.pi)
    assert m.log_prob(z[None])[0] - c["logj"][0] == pytest.approx(target, rel=1e-12)


def test_ucln_skygrid_model_gradient_matches_finite_differences():
    d, S, lowers, heights, lik, grid, m = _hcv_model(L=30)
    rng = np.random.default_rng(9)
    Z = np.stack([_ucln_point(m, heights, lowers, rng) for _ in range(2)])
    lp, Gd = m.log_prob_grad(Z)
    assert np.all(np.isfinite(lp)) and lik.calls == 1
    for k in list(range(0, m.dim, 4)) + list(range(m.slices["thetas"].start, m.slices["tau"].stop)) + \
            [m.slices["ucln_mean"].start, m.slices["ucln_stdev"].start, m.slices["height"].start]:
        e = np.zeros(m.dim); e[k] = 1e-6
        fd = (m.log_prob(Z + e) - m.log_prob(Z - e)) / 2e-6
        assert np.allclose(fd, Gd[:, k], rtol=1e-4, atol=1e-4), (k, fd, Gd[:, k])
    # large batches take the vectorised bincount path of the skygrid gradient
    Z8 = np.stack([_ucln_point(m, heights, lowers, rng) for _ in range(6)])
    lp8, G8 = m.log_prob_grad(Z8)
    lp1, G1 = m.log_prob_grad(Z8[4:5])
    assert np.allclose(lp8[4], lp1[0], rtol=1e-13) and np.allclose(G8[4], G1[0], rtol=1e-10, atol=1e-10)
    assert len(m.constrained_names()) == m.constrained_matrix(Z).shape[1]


@pytest.mark.gpu
def test_gpu_hcv_ucln_skygrid_model_block():
    """BASELINE config 5's model (HCV, GTR + W4, heterochronous, ucln clock, skygrid): the model block on the
    library equals the one on the oracle back end, and a short batched ADVI run raises the ELBO."""
    from phylostan_b200 import likelihood as lk
    d, S, lowers, heights, ora, grid, m_cpu = _hcv_model()
    with lk.TreeLikelihood(d["peel"], d["tipmask"], d["weights"], model="GTR", categories=4, rooted=True) as lik:
        m = advi.ClockModel(lik, "GTR", d["map"], lowers, clock="ucln", coalescent="skygrid", grid=grid)
        rng = np.random.default_rng(2)
        Z = np.stack([_ucln_point(m, heights, lowers, rng) for _ in range(4)])
        lp, Gd = m.log_prob_grad(Z)
        lp0, G0 = m_cpu.log_prob_grad(Z)
        assert np.max(np.abs(lp - lp0) / np.abs(lp0)) <= 1e-10
        assert np.max(np.abs(Gd - G0) / np.maximum(1.0, np.abs(G0))) <= 1e-7
        fit = advi.advi(m, iter=600, grad_samples=8, elbo_samples=50, eval_elbo=100, eta=0.1, seed=1, init=Z[0],
                        output_samples=100)
    assert fit.elbo_trace[-1][1] > fit.elbo_trace[0][1]
    mean = fit.mean()
    assert mean["height"] > lowers.max() and 1e-5 < mean["ucln_mean"] < 1e-1


def test_a_rejected_batch_is_retried_draw_by_draw():
    """The library answers a batch containing a draw with a non-finite likelihood with its domain error; the
    model block then isolates that draw (-inf, like a draw Stan rejects) and keeps the others."""
    class PhyloDomainError(Exception):
        pass

    class Picky(OracleLikelihood):
        def _each(self, blens, subst, freqs, rs, ps, want_grad):
            if np.any(blens > 50.0):
                self.calls += 1
                raise PhyloDomainError("log-likelihood is not finite")
            return super()._each(blens, subst, freqs, rs, ps, want_grad)

    prob = synth.make_problem(6, 40, 4, seed=5)
    lik = Picky(E.unrooted_swap(prob.peel), prob.tipmask, prob.weights, "GTR", 4)
    m = advi.UnrootedModel(lik, "GTR")
    Z = np.random.default_rng(0).normal(-1.0, 0.3, (3, m.dim))
    Z[1, m.slices["blens"].start] = 5.0                                  # exp(5) > 50
    lp, G = m.log_prob_grad(Z)
    assert np.isfinite(lp[0]) and np.isfinite(lp[2]) and lp[1] == -np.inf
    good = m.log_prob_grad(Z[[0, 2]])
    assert np.allclose(lp[[0, 2]], good[0], rtol=1e-13) and np.allclose(G[[0, 2]], good[1], rtol=1e-12)
    assert np.array_equal(m.log_prob(Z) == -np.inf, [False, True, False])


def _write_newick_and_fasta(tmp_path, S=8, L=300, seed=3):
    """A small simulated data set as the files `phylostan run -t ... -i ...` takes."""
    prob = synth.make_problem(S, L, 4, seed=seed)
    names = [f"t{k + 1}_{2000 + (k * 3) % 11}" for k in range(S)]
    sub = {k: names[k] for k in range(S)}
    for a, b, p in prob.peel:
        sub[p - 1] = "(%s:%.6f,%s:%.6f)" % (sub[a - 1], max(prob.blens[a - 1], 1e-4) * 50, sub[b - 1],
                                             max(prob.blens[b - 1], 1e-4) * 50)
    tree = tmp_path / "sim.tree"
    tree.write_text(sub[2 * S - 2] + ";\n")
    code = {1: "A", 2: "C", 4: "G", 8: "T", 15: "N"}
    fasta = tmp_path / "sim.fa"
    with open(fasta, "w") as f:
        for k in range(S):
            seq = "".join(code.get(int(m), "N") * int(w) for m, w in zip(prob.tipmask[k], prob.weights))
            f.write(f">{names[k]}\n{seq}\n")
    return tree, fasta


def test_command_line_reads_tree_and_alignment(tmp_path):
    """The file front end of the drivers (no GPU needed up to the encoders)."""
    tree, fasta = _write_newick_and_fasta(tmp_path)
    t = E.read_tree(str(tree))
    enc = E.encode(t, E.read_alignment(str(fasta)), rooted=True)
    assert enc.S == 8 and enc.peel.shape == (7, 3) and enc.tipmask.shape[0] == 8 and enc.weights.sum() > 300
    assert E.setup_dates(t, None, True) > 0 and E.get_lowers(t).shape == (15,)


@pytest.mark.gpu
@pytest.mark.parametrize("argv", [
    ["-m", "HKY", "-C", "4", "--iter", "300", "--grad_samples", "4"],
    ["-m", "GTR", "-C", "4", "--clock", "strict", "--heterochronous", "--iter", "300", "--grad_samples", "4", "-q", "fullrank"],
    ["-m", "JC69", "--clock", "ucln", "-c", "skygrid", "--grid", "5", "--cutoff", "30", "--iter", "200"],
    ["-m", "HKY", "-C", "4", "--clock", "uced", "-c", "skyride", "--heterochronous", "--iter", "200", "--grad_samples", "2"],
    ["-m", "HKY", "--clock", "strict", "-a", "nuts", "--iter", "60"],
    ["-m", "HKY", "--clock", "strict", "--heterochronous", "-a", "hmc", "--chains", "4", "--iter", "60"],
])
def test_gpu_command_line_end_to_end(tmp_path, argv):
    """python -m phylostan_b200.advi: tree + alignment files in, CSV of posterior draws out."""
    tree, fasta = _write_newick_and_fasta(tmp_path)
    out = tmp_path / "draws.csv"
    assert advi.main(["-t", str(tree), "-i", str(fasta), "-o", str(out), "--samples", "50"] + argv) == 0
    rows = out.read_text().strip().split("\n")
    header = rows[0].split(",")
    data = np.array([[float(x) for x in r.split(",")] for r in rows[1:]])
    assert data.shape[1] == len(header) and data.shape[0] >= 30 and np.all(np.isfinite(data))
    if "--clock" in argv:
        assert "height" in header and np.all(data[:, header.index("height")] > 0)
    else:
        assert "blens.1" in header and np.all(data[:, header.index("blens.1")] > 0)


def test_uced_skyride_model_is_the_stan_program_and_its_gradient():
    """uced clock + skyride coalescent: skyride_coalescent_log (generate_script.py:352-414) and the exponential
    clock prior (:1319-1322) restated literally for one draw; gradient against finite differences."""
    d, S, lowers, heights = flua_clock_problem("HCV")
    L = 30
    lik = OracleRooted(d["peel"], d["tipmask"][:, :L].copy(), d["weights"][:L].copy(), "JC69", 1)
    m = advi.ClockModel(lik, "JC69", d["map"], lowers, clock="uced", coalescent="skyride")
    assert m.dim == (S - 2) + (2 * S - 2) + 1 + 1 + (S - 1) + 1
    rng = np.random.default_rng(6)
    z = rng.normal(0, 0.1, m.dim)
    props = np.array([(heights[m.tr_node[j]] - m.tr_lo[j]) / (heights[m.tr_parent[j]] - m.tr_lo[j]) for j in range(S - 2)])
    props = np.clip(props, 1e-4, 1 - 1e-4)
    z[m.slices["props"]] += np.log(props) - np.log1p(-props)
    z[m.slices["height"]] += math.log(heights[m.root - S - 1] - m.lower_root)
    z[m.slices["substrates"]] += math.log(8e-4)
    z[m.slices["uced_mean"]] += math.log(8e-4)
    z[m.slices["thetas"]] += 5.0
    c = m.constrain(z[None])
    hs, sub, mean, pop, tau = list(c["heights"][0]), c["substrates"][0], c["uced_mean"][0], c["thetas"][0], c["tau"][0]
    mp = [[int(a), int(b)] for a, b in d["map"]]
    nodeCount = 2 * S - 1
    blens = np.zeros(2 * S - 2)
    for i in range(1, nodeCount):
        node, par = mp[i]
        blens[node - 1] = sub[node - 1] * (hs[par - S - 1] - (hs[node - S - 1] if node > S else lowers[node - 1]))
    # skyride_coalescent_log: times / childCounts indexed by pre-order row, as in the Stan text
    times = [hs[mp[i][0] - S - 1] if mp[i][0] > S else lowers[mp[i][0] - 1] for i in range(nodeCount)]
    child = [2 if mp[i][0] > S else 0 for i in range(nodeCount)]
    order = sorted(range(nodeCount), key=lambda k: times[k])
    logP, index, lineages, start = 0.0, 1, 0.0, times[order[0]]
    for k in order:
        finish = times[k]
        interval = finish - start
        if interval != 0.0:
            logP -= interval * (lineages * (lineages - 1.0)) / 2.0 / math.exp(pop[index - 1])
            if child[k] != 0:
                logP -= pop[index - 1]
                index += 1
        lineages += 1.0 if child[k] == 0 else -1.0
        start = finish
    I = S - 1
    gmrf = math.log(tau) * (I - 1.0) / 2.0 - sum((pop[i] - pop[i - 1]) ** 2 for i in range(1, I)) * tau / 2.0 \
        - (I - 1.0) / 2.0 * math.log(2.0 * math.pi)
    target = sum(math.log(1.0 / mean) - x / mean for x in sub) - 1000.0 * mean        # exponential(1/uced_mean)
    target += logP + gmrf + (0.001 - 1.0) * math.log(tau) - 0.001 * tau
    target += O.loglik_grad(lik.peel, lik.tipmask, lik.weights, O.JC69, blens, None, None, np.ones(1), np.ones(1),
                            rooted=True, want_grad=False).logp
    for i in range(1, nodeCount):
        if mp[i][0] > S:
            target += math.log(hs[mp[i][1] - S - 1] - lowers[mp[i][0] - 1])
    assert m.log_prob(z[None])[0] - c["logj"][0] == pytest.approx(target, rel=1e-12)
    Z = np.stack([z, z + rng.normal(0, 0.05, m.dim)])
    lp, G = m.log_prob_grad(Z)
    for k in list(range(0, m.dim, 5)) + [m.slices["uced_mean"].start, m.slices["tau"].start, m.slices["height"].start]:
        e = np.zeros(m.dim); e[k] = 1e-6
        fd = (m.log_prob(Z + e) - m.log_prob(Z - e)) / 2e-6
        assert np.allclose(fd, G[:, k], rtol=1e-4, atol=1e-4), (k, fd, G[:, k])


@pytest.mark.gpu
def test_gpu_clock_model_device_front_end_equals_host_chain_rule(monkeypatch):
    """ClockModel on a library handle runs heights -> blens, the likelihood and the reverse sweeps down to d/dprops in
    one phylo_b200_eval_ratios_batch call; PHYLO_B200_HOST_FRONT_END=1 keeps the chain rule in numpy.  Same numbers,
    and a draw whose proportions make the library reject it is -inf with a zero gradient on both paths."""
    from phylostan_b200 import likelihood as lk
    d, S, lowers, heights, ora, grid, m_cpu = _hcv_model()
    with lk.TreeLikelihood(d["peel"], d["tipmask"], d["weights"], model="GTR", categories=4, rooted=True) as lik:
        m = advi.ClockModel(lik, "GTR", d["map"], lowers, clock="ucln", coalescent="skygrid", grid=grid)
        assert m._device_front_end()
        rng = np.random.default_rng(12)
        Z = np.stack([_ucln_point(m, heights, lowers, rng) for _ in range(5)])
        lp_d, G_d = m.log_prob_grad(Z)
        lv_d = m.log_prob(Z)
        monkeypatch.setenv("PHYLO_B200_HOST_FRONT_END", "1")
        assert not m._device_front_end()
        lp_h, G_h = m.log_prob_grad(Z)
    assert np.max(np.abs(lp_d - lp_h) / np.abs(lp_h)) <= 1e-12
    assert np.max(np.abs(lv_d - lp_h) / np.abs(lp_h)) <= 1e-12
    assert np.max(np.abs(G_d - G_h) / np.maximum(1.0, np.abs(G_h))) <= 1e-8
