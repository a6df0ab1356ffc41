"""NUTS and lock-step multi-chain HMC over the model blocks of ``phylostan_b200.advi`` (phylostan's ``-a nuts`` / ``-a hmc``).

phylostan hands sampling to Stan (``sm.sampling(algorithm='NUTS')``, phylostan/phylostan.py:318-321), where
every leapfrog step is one value-and-gradient evaluation of the model block -- the call the GPU library
serves.  This module is the same sampler outside Stan, for the model blocks that exist here
(``UnrootedModel``, ``StrictClockModel``): one ``log_prob_grad`` call (one ``phylo_b200_eval``) per
leapfrog step.

The algorithm is Stan 2.19's (third-party, not under /root/reference; restated from its published
description -- Hoffman & Gelman 2014 for the doubling scheme and dual averaging, Betancourt 2017
"A Conceptual Introduction to Hamiltonian Monte Carlo" for multinomial sampling and the generalised
no-U-turn criterion, and the Stan reference manual's "HMC algorithm parameters" / "Automatic parameter
tuning" for the constants): diagonal Euclidean metric; trajectory doubling up to ``max_depth`` = 10 with
multinomial sampling biased towards the new subtree at the top level; U-turn criterion on the summed
momenta ``rho``; divergence when the energy error exceeds 1000; step size by dual averaging towards
``delta`` = 0.8 (gamma 0.05, t0 10, kappa 0.75, mu = log(10 eps)); metric from windowed variance estimates
(initial buffer 75, base window 25 doubling, final buffer 50, shrunk towards 1e-3).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

__all__ = ["NutsFit", "nuts", "HmcFit", "hmc"]


@dataclass
class NutsFit:
    draws: np.ndarray                  # [num_samples, n constrained]
    names: List[str]
    unconstrained: np.ndarray          # [num_samples, dim]
    lp: np.ndarray                     # log density (with Jacobian) of every draw
    stepsize: float
    inv_metric: np.ndarray
    accept_stat: np.ndarray
    treedepth: np.ndarray
    n_leapfrog: np.ndarray
    divergent: np.ndarray
    gradient_evaluations: int = 0

    def mean(self) -> Dict[str, float]:
        return dict(zip(self.names, self.draws.mean(axis=0)))


class _State:
    __slots__ = ("q", "p", "V", "g")

    def __init__(self, q, p, V, g):
        self.q, self.p, self.V, self.g = q, p, V, g

    def copy(self):
        return _State(self.q.copy(), self.p.copy(), self.V, self.g.copy())


def _lse(a, b):
    if a == -math.inf:
        return b
    if b == -math.inf:
        return a
    m = max(a, b)
    return m + math.log(math.exp(a - m) + math.exp(b - m))


class _Nuts:
    MAX_DELTA_H = 1000.0

    def __init__(self, model, rng, max_depth):
        self.m, self.rng, self.max_depth = model, rng, max_depth
        self.inv = np.ones(model.dim)
        self.eps = 1.0
        self.ngrad = 0

    # potential V = -log p and its gradient
    def potential(self, q):
        lp, g = self.m.log_prob_grad(q[None, :])
        self.ngrad += 1
        if not np.isfinite(lp[0]) or not np.all(np.isfinite(g[0])):
            return math.inf, np.zeros_like(q)
        return -float(lp[0]), -g[0]

    def H(self, z):
        return z.V + 0.5 * float(np.dot(z.p, self.inv * z.p))

    def leapfrog(self, z, eps):
        z.p -= 0.5 * eps * z.g
        z.q += eps * self.inv * z.p
        z.V, z.g = self.potential(z.q)
        z.p -= 0.5 * eps * z.g

    def sample_p(self):
        return self.rng.standard_normal(self.m.dim) / np.sqrt(self.inv)

    def init_stepsize(self, q, V, g):
        """Stan's heuristic: double or halve eps until the acceptance probability of one step crosses 0.8."""
        z = _State(q.copy(), self.sample_p(), V, g.copy())
        H0 = self.H(z)
        self.leapfrog(z, self.eps)
        h = self.H(z)
        dH = H0 - (h if np.isfinite(h) else math.inf)
        direction = 1 if dH > math.log(0.8) else -1
        for _ in range(100):
            z = _State(q.copy(), self.sample_p(), V, g.copy())
            H0 = self.H(z)
            self.leapfrog(z, self.eps)
            h = self.H(z)
            dH = H0 - (h if np.isfinite(h) else math.inf)
            if direction == 1 and not dH > math.log(0.8):
                break
            if direction == -1 and not dH < math.log(0.8):
                break
            self.eps = 2.0 * self.eps if direction == 1 else 0.5 * self.eps
            if self.eps > 1e7 or self.eps == 0.0:
                raise RuntimeError("step-size initialisation failed")

    def build_tree(self, depth, z, ctx, sign):
        """Returns (valid, z_propose, p_sharp_left, p_sharp_right, rho, log_sum_weight); advances z in place."""
        if depth == 0:
            self.leapfrog(z, sign * self.eps)
            ctx["n_leapfrog"] += 1
            h = self.H(z)
            if not np.isfinite(h):
                h = math.inf
            if h - ctx["H0"] > self.MAX_DELTA_H:
                ctx["divergent"] = True
            d = ctx["H0"] - h
            ctx["sum_metro"] += 1.0 if d > 0 else math.exp(d)
            ps = self.inv * z.p
            return (not ctx["divergent"]), z.copy(), ps, ps.copy(), z.p.copy(), d
        ok_l, prop_l, psl_l, _, rho_l, lw_l = self.build_tree(depth - 1, z, ctx, sign)
        if not ok_l:
            return False, prop_l, psl_l, psl_l, rho_l, lw_l
        ok_r, prop_r, _, psr_r, rho_r, lw_r = self.build_tree(depth - 1, z, ctx, sign)
        if not ok_r:
            return False, prop_l, psl_l, psr_r, rho_l + rho_r, _lse(lw_l, lw_r)
        lw = _lse(lw_l, lw_r)
        prop = prop_r if self.rng.uniform() < math.exp(lw_r - lw) else prop_l
        rho = rho_l + rho_r
        ok = float(np.dot(psl_l, rho)) > 0 and float(np.dot(psr_r, rho)) > 0
        return ok, prop, psl_l, psr_r, rho, lw

    def transition(self, q, V, g):
        z0 = _State(q.copy(), self.sample_p(), V, g.copy())
        ctx = {"H0": self.H(z0), "n_leapfrog": 0, "sum_metro": 0.0, "divergent": False}
        z_plus, z_minus, sample = z0.copy(), z0.copy(), z0.copy()
        ps_plus = ps_minus = self.inv * z0.p
        rho = z0.p.copy()
        lw = 0.0
        depth = 0
        while depth < self.max_depth:
            if self.rng.uniform() > 0.5:
                ok, prop, _, ps_plus, rho_sub, lw_sub = self.build_tree(depth, z_plus, ctx, +1)
            else:
                ok, prop, _, ps_minus, rho_sub, lw_sub = self.build_tree(depth, z_minus, ctx, -1)
            if not ok:
                break
            depth += 1
            if lw_sub > lw or self.rng.uniform() < math.exp(lw_sub - lw):
                sample = prop
            lw = _lse(lw, lw_sub)
            rho = rho + rho_sub
            if not (float(np.dot(ps_plus, rho)) > 0 and float(np.dot(ps_minus, rho)) > 0):
                break
        accept = ctx["sum_metro"] / max(ctx["n_leapfrog"], 1)
        return sample, accept, depth, ctx["n_leapfrog"], ctx["divergent"]


class _DualAveraging:
    def __init__(self, delta=0.8, gamma=0.05, t0=10.0, kappa=0.75):
        self.delta, self.gamma, self.t0, self.kappa = delta, gamma, t0, kappa
        self.restart(1.0)

    def restart(self, eps):
        self.mu = math.log(10.0 * eps)
        self.counter, self.s_bar, self.x_bar = 0, 0.0, 0.0

    def learn(self, accept):
        self.counter += 1
        accept = min(1.0, accept)
        w = 1.0 / (self.counter + self.t0)
        self.s_bar = (1.0 - w) * self.s_bar + w * (self.delta - accept)
        x = self.mu - self.s_bar * math.sqrt(self.counter) / self.gamma
        x_eta = self.counter ** (-self.kappa)
        self.x_bar = (1.0 - x_eta) * self.x_bar + x_eta * x
        return math.exp(x)

    def final(self):
        return math.exp(self.x_bar)


def _windows(num_warmup, init_buffer=75, term_buffer=50, base_window=25):
    """End iterations (exclusive) of the metric-adaptation windows, Stan's windowed_adaptation."""
    if num_warmup < 20:
        return [], num_warmup
    if init_buffer + base_window + term_buffer > num_warmup:
        init_buffer, term_buffer = int(0.15 * num_warmup), int(0.1 * num_warmup)
        base_window = num_warmup - init_buffer - term_buffer
    ends, start, size = [], init_buffer, base_window
    last = num_warmup - term_buffer
    while start < last:
        end = start + size
        if end + 2 * size > last:          # the next window would not fit: stretch this one
            end = last
        ends.append(end)
        start, size = end, 2 * size
    return ends, init_buffer


def nuts(model, *, num_warmup: int = 1000, num_samples: int = 1000, seed: int = 1, init="random", max_depth: int = 10,
         delta: float = 0.8, stepsize: float = 1.0, verbose: bool = False) -> NutsFit:
    """One NUTS chain.  ``init``: "random" (uniform(-2, 2) unconstrained), "zero" or an unconstrained vector."""
    rng = np.random.default_rng(seed)
    if isinstance(init, str):
        q = rng.uniform(-2.0, 2.0, model.dim) if init == "random" else np.zeros(model.dim)
    else:
        q = np.asarray(init, dtype=np.float64).copy()
    S = _Nuts(model, rng, max_depth)
    S.eps = stepsize
    V, g = S.potential(q)
    if not np.isfinite(V):
        raise ValueError("the initial point has zero density")
    S.init_stepsize(q, V, g)
    da = _DualAveraging(delta)
    da.restart(S.eps)
    ends, init_buffer = _windows(num_warmup)
    win_n, win_mean, win_m2 = 0, np.zeros(model.dim), np.zeros(model.dim)
    total = num_warmup + num_samples
    out_q = np.empty((num_samples, model.dim))
    out_lp, acc, td, nl, dv = (np.empty(num_samples), np.empty(num_samples), np.empty(num_samples, dtype=int),
                               np.empty(num_samples, dtype=int), np.zeros(num_samples, dtype=bool))
    for it in range(total):
        z, a, depth, nleap, div = S.transition(q, V, g)
        q, V, g = z.q, z.V, z.g
        if it < num_warmup:
            S.eps = da.learn(a)
            if ends and init_buffer <= it < ends[-1]:
                win_n += 1                                              # Welford
                dlt = q - win_mean
                win_mean += dlt / win_n
                win_m2 += dlt * (q - win_mean)
                if it + 1 in ends:
                    var = win_m2 / max(win_n - 1, 1)
                    S.inv = (win_n / (win_n + 5.0)) * var + 1e-3 * (5.0 / (win_n + 5.0))
                    win_n, win_mean, win_m2 = 0, np.zeros(model.dim), np.zeros(model.dim)
                    S.init_stepsize(q, V, g)
                    da.restart(S.eps)
            if it + 1 == num_warmup:
                S.eps = da.final()
        else:
            k = it - num_warmup
            out_q[k], out_lp[k], acc[k], td[k], nl[k], dv[k] = q, -V, a, depth, nleap, div
        if verbose and (it + 1) % max(total // 10, 1) == 0:
            print(f"  iteration {it + 1}/{total}  eps {S.eps:.4g}  treedepth {depth}  accept {a:.2f}")
    return NutsFit(model.constrained_matrix(out_q), model.constrained_names(), out_q, out_lp, float(S.eps), S.inv.copy(),
                   acc, td, nl, dv, gradient_evaluations=S.ngrad)


# ---------------------------------------------------------------------------------------------------
# static HMC, many chains in lock step (one batched library call per leapfrog step)
# ---------------------------------------------------------------------------------------------------
@dataclass
class HmcFit:
    draws: np.ndarray                  # [chains, num_samples, n constrained]
    names: List[str]
    unconstrained: np.ndarray          # [chains, num_samples, dim]
    lp: np.ndarray                     # [chains, num_samples]
    stepsize: float
    n_leapfrog: int
    inv_metric: np.ndarray
    accept_stat: np.ndarray            # [chains, num_samples]
    gradient_calls: int = 0            # batched library calls (each evaluates every chain)

    def mean(self) -> Dict[str, float]:
        return dict(zip(self.names, self.draws.reshape(-1, self.draws.shape[-1]).mean(axis=0)))


def hmc(model, *, chains: int = 8, num_warmup: int = 1000, num_samples: int = 1000, int_time: float = 2.0 * math.pi,
        seed: int = 1, init="random", delta: float = 0.8, stepsize: float = 1.0, max_leapfrog: int = 256,
        verbose: bool = False) -> HmcFit:
    """Static HMC (phylostan's ``-a hmc``; Stan's ``static`` engine: integration time ``int_time`` = 2 pi,
    at most L = int_time / eps leapfrog steps -- the count is jittered uniformly in 1..L -- diagonal metric) for ``chains`` chains advanced in LOCK STEP: every
    leapfrog step is one batched ``log_prob_grad`` call, i.e. one ``phylo_b200_eval_batch`` for all chains --
    what Stan's one-process-per-chain sampling (``--chains``, phylostan/phylostan.py:319-321) cannot do.
    Step size (dual averaging on the mean acceptance probability) and metric (windowed variance pooled
    over the chains) are shared by the chains.  ``init``: "random", "zero" or an unconstrained vector /
    [chains, dim] array."""
    rng = np.random.default_rng(seed)
    d = model.dim
    if isinstance(init, str):
        q = rng.uniform(-2.0, 2.0, (chains, d)) if init == "random" else np.zeros((chains, d))
    else:
        q = np.broadcast_to(np.asarray(init, dtype=np.float64), (chains, d)).copy()
    calls = 0

    def potential(qq):
        nonlocal calls
        lp, g = model.log_prob_grad(qq)
        calls += 1
        bad = ~np.isfinite(lp) | ~np.all(np.isfinite(g), axis=1)
        return np.where(bad, np.inf, -lp), np.where(bad[:, None], 0.0, -g)

    V, g = potential(q)
    if not np.all(np.isfinite(V)):
        raise ValueError("an initial point has zero density")
    inv = np.ones(d)
    eps = float(stepsize)

    def trajectory(q, V, g, eps, L):
        p = rng.standard_normal((chains, d)) / np.sqrt(inv)
        H0 = V + 0.5 * (p * p * inv).sum(axis=1)
        qn, pn, Vn, gn = q.copy(), p, V, g
        for _ in range(L):
            pn = pn - 0.5 * eps * gn
            qn = qn + eps * inv * pn
            Vn, gn = potential(qn)
            pn = pn - 0.5 * eps * gn
        H1 = Vn + 0.5 * (pn * pn * inv).sum(axis=1)
        with np.errstate(over="ignore", invalid="ignore"):
            a = np.where(np.isfinite(H1), np.minimum(1.0, np.exp(H0 - H1)), 0.0)
        acc = rng.uniform(size=chains) < a
        return (np.where(acc[:, None], qn, q), np.where(acc, Vn, V), np.where(acc[:, None], gn, g), a)

    def find_eps(eps):
        """Stan's heuristic on the mean one-step acceptance probability of the chains."""
        def one(e):
            return float(trajectory(q, V, g, e, 1)[3].mean())
        direction = 1 if one(eps) > 0.8 else -1
        for _ in range(60):
            a = one(eps)
            if (direction == 1 and not a > 0.8) or (direction == -1 and not a < 0.8):
                break
            eps = 2.0 * eps if direction == 1 else 0.5 * eps
        return eps

    eps = find_eps(eps)
    da = _DualAveraging(delta)
    da.restart(eps)
    ends, init_buffer = _windows(num_warmup)
    win_n, win_mean, win_m2 = 0, np.zeros(d), np.zeros(d)
    out_q = np.empty((chains, num_samples, d))
    out_lp, acc = np.empty((chains, num_samples)), np.empty((chains, num_samples))
    L = 1
    for it in range(num_warmup + num_samples):
        L = int(min(max(1, math.floor(int_time / eps)), max_leapfrog))
        # the number of steps is drawn uniformly from 1..L: a fixed integration time resonates with the
        # period of a well-adapted Gaussian-like posterior (2 pi is exactly one period)
        q, V, g, a = trajectory(q, V, g, eps, int(rng.integers(1, L + 1)))
        if it < num_warmup:
            eps = da.learn(float(a.mean()))
            if ends and init_buffer <= it < ends[-1]:
                for cq in q:                                            # Welford, pooled over the chains
                    win_n += 1
                    dlt = cq - win_mean
                    win_mean += dlt / win_n
                    win_m2 += dlt * (cq - win_mean)
                if it + 1 in ends:
                    var = win_m2 / max(win_n - 1, 1)
                    inv = (win_n / (win_n + 5.0)) * var + 1e-3 * (5.0 / (win_n + 5.0))
                    win_n, win_mean, win_m2 = 0, np.zeros(d), np.zeros(d)
                    eps = find_eps(eps)
                    da.restart(eps)
            if it + 1 == num_warmup:
                eps = da.final()
        else:
            k = it - num_warmup
            out_q[:, k], out_lp[:, k], acc[:, k] = q, -V, a
        if verbose and (it + 1) % max((num_warmup + num_samples) // 10, 1) == 0:
            print(f"  iteration {it + 1}  eps {eps:.4g}  L {L}  accept {a.mean():.2f}")
    draws = model.constrained_matrix(out_q.reshape(-1, d))
    return HmcFit(draws.reshape(chains, num_samples, -1), model.constrained_names(), out_q, out_lp, float(eps), L,
                  inv.copy(), acc, gradient_calls=calls)
