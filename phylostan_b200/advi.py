"""Batched ADVI (mean-field or full-rank) over the GPU likelihood.

Stan's ADVI evaluates ``log_prob`` one draw at a time (``grad_samples`` value+gradient calls per
iteration, ``elbo_samples`` value-only calls per ELBO estimate; phylostan/phylostan.py:47-50,311-313).
Here all draws of an iteration go through ONE ``phylo_b200_eval_batch`` call, which is the only way
the batch parallelism of the likelihood library (BASELINE config 3) reaches a user.

Scope: model blocks of the programs phylostan generates (phylostan/generate_script.py:1186-1457).
``UnrootedModel``: an unconstrained, unrooted tree (``clock is None``; tests/golden/DS1-GTR-W4-external.stan):
``wshape`` (Weibull categories), ``blens``, ``rates``/``kappa`` + ``freqs``; priors ``wshape ~ exponential(1)``,
``blens ~ exponential(10)``, ``rates ~ dirichlet(rates_alpha)``, ``freqs ~ dirichlet(frequencies_alpha)``,
``kappa ~ lognormal(1, 1.25)``.  ``ClockModel``: a time tree (tips dated or not) with ratio-transformed node
heights, a strict or uncorrelated (lognormal / exponential) clock and a constant-size, skyride or skygrid
coalescent; ``StrictClockModel`` = strict + constant, the fluA quick start
(tests/golden/fluA-HKY-W4-external.stan); ucln + skygrid is BASELINE config 5's program (HCV).  The
autocorrelated clocks and the birth-death prior stay in Stan and use the external-function route
(INTEGRATION.md).

The algorithm follows Stan 2.19's ``stan::variational::advi`` with the ``normal_meanfield`` (zeta = mu +
exp(omega) * eta) or ``normal_fullrank`` (zeta = mu + L eta, L lower triangular) family
(third-party, not under /root/reference; restated from its published description: Kucukelbir et al.
2017, "Automatic Differentiation Variational Inference", Alg. 1 and the adaptive step-size sequence
of section 2.6 / Stan reference manual "ADVI algorithm"):  unconstrained parameters zeta = mu +
exp(omega) * eta, eta ~ N(0, I); gradient of the ELBO by the reparameterisation trick; step size
eta_k = eta * k^(-1/2) / (tau + sqrt(s_k)), s_k = alpha g_k^2 + (1 - alpha) s_{k-1}, alpha = 0.1,
tau = 1; eta adapted over (100, 10, 1, 0.1, 0.01) with 50 iterations each; convergence when the mean
or the median of the relative ELBO changes in a circular buffer falls below ``tol_rel_obj``.
Constraining transforms and their log-Jacobians are Stan's (lower bound: x = a + exp(u); simplex:
stick breaking with the log(K - k) offset).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

__all__ = ["UnrootedModel", "ClockModel", "StrictClockModel", "MeanFieldFit", "advi", "advi_meanfield", "simplex_constrain", "simplex_adjoint", "weibull_rates"]


# ---------------------------------------------------------------------------------------------------
# transforms (batched over the leading axis)
# ---------------------------------------------------------------------------------------------------
def simplex_constrain(y: np.ndarray) -> Tuple[np.ndarray, np.ndarray, Tuple[np.ndarray, np.ndarray]]:
    """Stan's stick-breaking transform R^(K-1) -> K-simplex.  Returns (x [B,K], log|J| [B], cache)."""
    y = np.asarray(y, dtype=np.float64)
    B, Km1 = y.shape
    K = Km1 + 1
    z = 1.0 / (1.0 + np.exp(-(y - np.log(K - 1 - np.arange(Km1)))))
    x = np.empty((B, K))
    rem = np.empty((B, Km1))
    r = np.ones(B)
    for k in range(Km1):
        rem[:, k] = r
        x[:, k] = r * z[:, k]
        r = r * (1.0 - z[:, k])
    x[:, K - 1] = r
    logj = np.sum(np.log(z) + np.log1p(-z) + np.log(rem), axis=1)
    return x, logj, (z, rem)


def simplex_adjoint(gx: np.ndarray, cache) -> np.ndarray:
    """d/dy of  f(x(y)) + log|J|(y)  given gx = df/dx  (reverse sweep of ``simplex_constrain``)."""
    z, rem = cache
    B, Km1 = z.shape
    gy = np.empty((B, Km1))
    rbar = gx[:, Km1].copy()                      # adjoint of the remaining stick after step k
    for k in range(Km1 - 1, -1, -1):
        zk, rk = z[:, k], rem[:, k]
        zbar = (gx[:, k] - rbar) * rk + 1.0 / zk - 1.0 / (1.0 - zk)
        rbar = gx[:, k] * zk + rbar * (1.0 - zk) + 1.0 / rk
        gy[:, k] = zbar * zk * (1.0 - zk)
    return gy


def weibull_rates(wshape: np.ndarray, C: int) -> Tuple[np.ndarray, np.ndarray]:
    """Discretised Weibull site rates (generate_script.py:267-278) and d rs / d wshape; [B,C] each."""
    w = np.asarray(wshape, dtype=np.float64).reshape(-1, 1)
    q = -np.log(1.0 - (2.0 * np.arange(C) + 1.0) / (2.0 * C))
    a = q[None, :] ** (1.0 / w)
    da = a * np.log(q)[None, :] * (-1.0 / (w * w))
    m = a.mean(axis=1, keepdims=True)
    rs = a / m
    drs = (da - rs * da.mean(axis=1, keepdims=True)) / m
    return rs, drs


# ---------------------------------------------------------------------------------------------------
# the model block
# ---------------------------------------------------------------------------------------------------
class _ModelBase:
    """Pieces shared by the model blocks: parameter layout in the Stan program's order, the site model
    (``wshape`` -> Weibull ``rs``), the substitution-model parameters with their priors
    (generate_script.py:1416-1457), and the reverse sweep through the constraining transforms."""

    WSHAPE_LOWER = 0.1           # real<lower=0.1> wshape   (generate_script.py:1212)

    def __init__(self, lik, model, rates_alpha, freqs_alpha):
        if model not in ("JC69", "HKY", "GTR"):
            raise ValueError("model must be JC69, HKY or GTR")
        self.lik, self.model, self.C, self.bcount = lik, model, int(lik.C), int(lik.bcount)
        self.rates_alpha = np.ones(6) if rates_alpha is None else np.asarray(rates_alpha, dtype=np.float64)
        self.freqs_alpha = np.ones(4) if freqs_alpha is None else np.asarray(freqs_alpha, dtype=np.float64)
        self.slices: Dict[str, slice] = {}
        self.dim = 0

    def _layout(self, blocks) -> None:
        o = 0
        for name, n in blocks:
            if n:
                self.slices[name] = slice(o, o + n)
                o += n
        self.dim = o

    def _subst_blocks(self):
        return (("rates", 5 if self.model == "GTR" else 0), ("kappa", 1 if self.model == "HKY" else 0),
                ("freqs", 3 if self.model != "JC69" else 0))

    def _subst_names(self) -> List[str]:
        names = [f"rates.{i + 1}" for i in range(6)] if self.model == "GTR" else []
        names += ["kappa"] if self.model == "HKY" else []
        return names + ([f"freqs.{i + 1}" for i in range(4)] if self.model != "JC69" else [])

    # -- constrain: shared parameters
    def _constrain_common(self, Z, out) -> None:
        if "wshape" in self.slices:
            u = Z[:, self.slices["wshape"]][:, 0]
            out["wshape"] = self.WSHAPE_LOWER + np.exp(u)
            out["logj"] += u
        if self.model == "GTR":
            out["rates"], lj, out["_rates"] = simplex_constrain(Z[:, self.slices["rates"]])
            out["logj"] += lj
        if self.model == "HKY":
            u = Z[:, self.slices["kappa"]][:, 0]
            out["kappa"] = np.exp(u)
            out["logj"] += u
        if self.model != "JC69":
            out["freqs"], lj, out["_freqs"] = simplex_constrain(Z[:, self.slices["freqs"]])
            out["logj"] += lj

    def _ok_common(self, c, rs) -> np.ndarray:
        ok = np.isfinite(c["logj"]) & np.all(np.isfinite(rs), axis=1) & np.all(rs > 0, axis=1)
        for k in ("rates", "freqs"):
            if k in c:
                ok &= np.all(c[k] > 0, axis=1)
        if "kappa" in c:
            ok &= np.isfinite(c["kappa"]) & (c["kappa"] > 0)
        return ok

    def _site_model(self, c, B):
        if self.C > 1:
            rs, drs = weibull_rates(c["wshape"], self.C)
        else:
            rs, drs = np.ones((B, 1)), np.zeros((B, 1))
        return rs, drs, np.full((B, self.C), 1.0 / self.C)

    def _subst_arg(self, c):
        return c["rates"] if self.model == "GTR" else c["kappa"][:, None] if self.model == "HKY" else None

    @staticmethod
    def _take(c, idx):
        return {k: (v[idx] if isinstance(v, np.ndarray) else tuple(a[idx] for a in v)) for k, v in c.items()}

    # -- `~` statements of the shared parameters; constants dropped exactly as Stan does
    def _prior_common(self, sub) -> np.ndarray:
        n = sub["logj"].shape[0]
        prior = np.zeros(n)
        if self.C > 1:
            prior += -sub["wshape"]                                     # wshape ~ exponential(1.0)
        if self.model == "GTR":
            prior += ((self.rates_alpha - 1.0) * np.log(sub["rates"])).sum(axis=1)
        if self.model == "HKY":
            lk = np.log(sub["kappa"])
            prior += -lk - (lk - 1.0) ** 2 / (2.0 * 1.25 ** 2)          # kappa ~ lognormal(1.0, 1.25)
        if self.model != "JC69":
            prior += ((self.freqs_alpha - 1.0) * np.log(sub["freqs"])).sum(axis=1)
        return prior

    # -- gradient w.r.t. the unconstrained shared parameters (likelihood + prior + log-Jacobian)
    def _grad_common(self, g, sub, vg, drs) -> None:
        n = g.shape[0]
        if self.C > 1:                                                  # x = 0.1 + exp(u)
            gw = (np.reshape(vg.grad_rs, (n, self.C)) * drs).sum(axis=1) - 1.0
            g[:, self.slices["wshape"]] = (gw * (sub["wshape"] - self.WSHAPE_LOWER) + 1.0)[:, None]
        if self.model == "GTR":
            gr = np.reshape(vg.grad_subst, (n, 6)) + (self.rates_alpha - 1.0) / sub["rates"]
            g[:, self.slices["rates"]] = simplex_adjoint(gr, sub["_rates"])
        if self.model == "HKY":
            k = sub["kappa"]
            gk = np.reshape(vg.grad_subst, (n,)) - 1.0 / k - (np.log(k) - 1.0) / (1.25 ** 2 * k)
            g[:, self.slices["kappa"]] = (gk * k + 1.0)[:, None]
        if self.model != "JC69":
            gf = np.reshape(vg.grad_freqs, (n, 4)) + (self.freqs_alpha - 1.0) / sub["freqs"]
            g[:, self.slices["freqs"]] = simplex_adjoint(gf, sub["_freqs"])

    def _likelihood(self, args, want_grad):
        """One batched library call.  Rejected draws (parameters out of domain, a non-finite likelihood -- the
        ``std::domain_error`` case of the Stan shim) get -inf and a zero gradient, which is what Stan does with a
        rejected draw: ``phylo_b200_eval_batch_status`` marks them inside the one call; a back end without it (the
        CPU stand-ins of the tests, ``ShardedLikelihood``) is retried draw by draw."""
        from .likelihood import ValueGrad
        if hasattr(self.lik, "value_grad_masked"):  # the library marks rejected draws itself (eval_batch_status)
            vg, _ = self.lik.value_grad_masked(*args, want_grad=want_grad)
            return np.atleast_1d(vg.log_P), (vg if want_grad else None)
        try:
            if want_grad:
                vg = self.lik.value_grad(*args)
                return np.atleast_1d(vg.log_P), vg
            return np.atleast_1d(self.lik.loglik(*args)), None
        except Exception as e:                                   # PhyloDomainError of the library (or a back end's own)
            if type(e).__name__ != "PhyloDomainError" or args[0].shape[0] == 1:
                if type(e).__name__ != "PhyloDomainError":
                    raise
                n = args[0].shape[0]
                zero = ValueGrad(np.full(n, -np.inf), np.zeros((n, self.bcount)), np.zeros((n, max(self.lik.nsubst, 0))),
                                 np.zeros((n, 4)), np.zeros((n, self.C)), np.zeros((n, self.C)))
                return zero.log_P, (zero if want_grad else None)
        parts = [self._likelihood(tuple(None if a is None else a[i:i + 1] for a in args), want_grad)
                 for i in range(args[0].shape[0])]
        ll = np.concatenate([p[0] for p in parts])
        if not want_grad:
            return ll, None
        cat = lambda name: np.concatenate([np.reshape(getattr(p[1], name), (1, -1)) for p in parts])
        return ll, ValueGrad(ll, cat("grad_blens"), cat("grad_subst"), cat("grad_freqs"), cat("grad_rs"), cat("grad_ps"))

    def log_prob(self, Z: np.ndarray) -> np.ndarray:
        """Value only, [B]; draws whose constrained values are not finite get -inf (Stan drops them)."""
        return self.log_prob_grad(Z, want_grad=False)[0]


class UnrootedModel(_ModelBase):
    """Jacobian-adjusted log density of the unrooted-tree program on Stan's unconstrained space.

    ``lik`` is a ``phylostan_b200.likelihood.TreeLikelihood`` created with ``rooted=False`` (anything
    with the same ``value_grad`` / ``loglik`` / ``bcount`` / ``C`` / ``nsubst`` surface works; the CPU
    tests use that to check the model block without a GPU).  Parameter order is the Stan program's:
    ``wshape`` (when C > 1), ``blens``, then ``rates`` (GTR) or ``kappa`` (HKY), then ``freqs``.
    """

    def __init__(self, lik, model: str = "GTR", rates_alpha=None, freqs_alpha=None):
        super().__init__(lik, model, rates_alpha, freqs_alpha)
        self._layout((("wshape", 1 if self.C > 1 else 0), ("blens", self.bcount)) + self._subst_blocks())

    # -- names of the constrained quantities, Stan CSV style
    def constrained_names(self) -> List[str]:
        return (["wshape"] if self.C > 1 else []) + [f"blens.{i + 1}" for i in range(self.bcount)] + self._subst_names()

    def constrain(self, Z: np.ndarray) -> Dict[str, np.ndarray]:
        """Unconstrained [B, dim] -> dict of constrained arrays plus ``logj`` and transform caches."""
        Z = np.atleast_2d(np.asarray(Z, dtype=np.float64))
        out: Dict[str, np.ndarray] = {"logj": np.zeros(Z.shape[0])}
        u = Z[:, self.slices["blens"]]
        out["blens"] = np.exp(u)
        out["logj"] += u.sum(axis=1)
        self._constrain_common(Z, out)
        return out

    def constrained_matrix(self, Z: np.ndarray) -> np.ndarray:
        c = self.constrain(Z)
        cols = [c["wshape"][:, None]] if self.C > 1 else []
        cols.append(c["blens"])
        for k in ("rates", "kappa", "freqs"):
            if k in c:
                cols.append(c[k] if c[k].ndim == 2 else c[k][:, None])
        return np.concatenate(cols, axis=1)

    def log_prob_grad(self, Z: np.ndarray, want_grad: bool = True) -> Tuple[np.ndarray, Optional[np.ndarray]]:
        Z = np.atleast_2d(np.asarray(Z, dtype=np.float64))
        B = Z.shape[0]
        with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
            c = self.constrain(Z)
            rs, drs, ps = self._site_model(c, B)
            ok = self._ok_common(c, rs) & np.all(np.isfinite(c["blens"]), axis=1) & np.all(c["blens"] < 1e6, axis=1)
        lp = np.full(B, -np.inf)
        G = np.zeros((B, self.dim)) if want_grad else None
        if not ok.any():
            return lp, G
        idx = slice(None) if ok.all() else np.nonzero(ok)[0]                     # the usual case: no copy
        sub = c if isinstance(idx, slice) else self._take(c, idx)
        rs_, drs_, ps_ = rs[idx], drs[idx], ps[idx]
        args = (sub["blens"], self._subst_arg(sub), sub.get("freqs"), rs_, ps_)
        ll, vg = self._likelihood(args, want_grad)
        prior = self._prior_common(sub) - 10.0 * sub["blens"].sum(axis=1)          # blens ~ exponential(10)
        lp[idx] = ll + prior + sub["logj"]
        if not want_grad:
            return lp, None
        n = ll.shape[0]
        g = np.zeros((n, self.dim))
        gb = np.reshape(vg.grad_blens, (n, self.bcount)) - 10.0
        g[:, self.slices["blens"]] = gb * sub["blens"] + 1.0
        self._grad_common(g, sub, vg, drs_)
        G[idx] = g
        return lp, G


class ClockModel(_ModelBase):
    """The programs phylostan generates for a time tree (generate_script.py:1290-1415): ratio-transformed
    node heights (:711-752), a clock, a coalescent prior, and the likelihood of the branch lengths
    ``rate * time`` (:660-679).

    ``clock``: "strict" (one ``rate``, ``rate ~ exponential(1000)``), "ucln" (one rate per branch,
    ``substrates ~ lognormal(log(ucln_mean) - ucln_stdev^2/2, ucln_stdev)``, ``ucln_mean ~ exponential(1000)``,
    ``ucln_stdev ~ gamma(0.5396, 2.6184)``; :1311-1318) or "uced" (``substrates ~ exponential(1/uced_mean)``,
    ``uced_mean ~ exponential(1000)``; :1319-1322).  ``coalescent``: "constant" (``theta ~ 1/x``,
    ``constant_coalescent_log`` :285-349), "skyride" (one log population size per coalescent interval,
    ``skyride_coalescent_log`` :352-414) or "skygrid" (log population sizes ``thetas[G]`` on the grid
    ``linspace(0, cutoff, grid)[1:]``, ``skygrid_coalescent_log`` :422-520), the last two with
    ``thetas ~ gmrf(tau)``, ``tau ~ gamma(0.001, 0.001)``.  Parameters in the Stan program's order: ``wshape`` (C > 1),
    ``props[S-2]``, clock parameters, root ``height``, coalescent parameters, ``rates``/``kappa``, ``freqs``.
    ``map_`` is the pre-order [node, parent] table and ``lowers`` the per-node lower bounds (tip dates) of
    ``phylostan_b200.encode`` (utils.py:84-104); ``lowers=None`` means contemporaneous tips.  Needs a
    ``rooted=True`` likelihood handle.
    """

    def __init__(self, lik, model: str, map_, lowers=None, lower_root: Optional[float] = None, rates_alpha=None,
                 freqs_alpha=None, clock: str = "strict", coalescent: str = "constant", grid=None):
        super().__init__(lik, model, rates_alpha, freqs_alpha)
        if clock not in ("strict", "ucln", "uced") or coalescent not in ("constant", "skyride", "skygrid"):
            raise ValueError("clock must be strict, ucln or uced; coalescent constant, skyride or skygrid")
        self.clock, self.coalescent = clock, coalescent
        m = np.asarray(map_, dtype=np.int64)
        self.S = S = (m.shape[0] + 1) // 2
        if m.shape != (2 * S - 1, 2) or self.bcount != 2 * S - 2:
            raise ValueError("map must be [2S-1, 2] and the likelihood handle rooted")
        self.nn = nn = 2 * S - 1
        self.lowers = np.zeros(nn) if lowers is None else np.asarray(lowers, dtype=np.float64)
        self.lower_root = float(self.lowers.max() if lower_root is None else max(lower_root, self.lowers.max()))
        self.root = int(m[0, 0])
        self.node = m[1:, 0] - 1                                 # 0-based node of every non-root pre-order row
        self.parent_h = m[1:, 1] - S - 1                         # index of its parent in heights[]
        self.internal = m[1:, 0] > S
        self.node_h = np.where(self.internal, m[1:, 0] - S - 1, 0)
        # internal non-root nodes in pre-order: (heights index, parent heights index, lower bound); prop j = position
        rows = np.nonzero(self.internal)[0]
        self.tr_node, self.tr_parent, self.tr_lo = self.node_h[rows], self.parent_h[rows], self.lowers[self.node[rows]]
        self.map32 = np.ascontiguousarray(m, dtype=np.int32)
        self.lowers_or_none = None if lowers is None else self.lowers
        # scatter of the reverse sweep d blens -> d heights: +1 at the parent, -1 at an internal node itself
        from scipy import sparse
        nb = 2 * S - 2
        rows = np.concatenate([np.arange(nb), np.nonzero(self.internal)[0]])
        cols = np.concatenate([self.parent_h, self.node_h[self.internal]])
        vals = np.concatenate([np.ones(nb), -np.ones(int(self.internal.sum()))])
        sc = sparse.csr_matrix((vals, (rows, cols)), shape=(nb, S - 1))
        self.scatter_blens = sc.toarray() if S <= 2000 else sc.T.tocsr()
        if coalescent == "skygrid":
            if grid is None or len(grid) < 2:
                raise ValueError("skygrid needs the grid points linspace(0, cutoff, grid)[1:]")
            self.grid = np.asarray(grid, dtype=np.float64)
            self.G = self.grid.size
        elif coalescent == "skyride":
            self.G = S - 1                                        # data['I']: one log population size per coalescent interval
        clock_blocks = {"strict": (("rate", 1),), "ucln": (("substrates", nb), ("ucln_mean", 1), ("ucln_stdev", 1)),
                        "uced": (("substrates", nb), ("uced_mean", 1))}[clock]
        coal_blocks = (("theta", 1),) if coalescent == "constant" else (("thetas", self.G), ("tau", 1))
        self._layout((("wshape", 1 if self.C > 1 else 0), ("props", S - 2)) + clock_blocks + (("height", 1),)
                     + coal_blocks + self._subst_blocks())

    def _scalar_names(self):
        clock = {"strict": ["rate"], "ucln": ["ucln_mean", "ucln_stdev"], "uced": ["uced_mean"]}[self.clock]
        coal = ["theta"] if self.coalescent == "constant" else ["tau"]
        return clock, coal

    def constrained_names(self) -> List[str]:
        names = (["wshape"] if self.C > 1 else []) + [f"props.{i + 1}" for i in range(self.S - 2)]
        if self.clock == "strict":
            names += ["rate"]
        else:
            names += [f"substrates.{i + 1}" for i in range(self.bcount)] + self._scalar_names()[0]
        names += ["height"]
        names += ["theta"] if self.coalescent == "constant" else [f"thetas.{i + 1}" for i in range(self.G)] + ["tau"]
        return names + self._subst_names() + [f"heights.{i + 1}" for i in range(self.S - 1)]

    def constrain(self, Z: np.ndarray) -> Dict[str, np.ndarray]:
        Z = np.atleast_2d(np.asarray(Z, dtype=np.float64))
        B = Z.shape[0]
        out: Dict[str, np.ndarray] = {"logj": np.zeros(B)}
        u = Z[:, self.slices["props"]]
        out["props"] = 1.0 / (1.0 + np.exp(-u))
        out["logj"] += -(np.logaddexp(0.0, u) + np.logaddexp(0.0, -u)).sum(axis=1)      # log p(1-p)
        clock, coal = self._scalar_names()
        for name in clock + ["height"] + coal:
            u = Z[:, self.slices[name]][:, 0]
            out[name] = (self.lower_root if name == "height" else 0.0) + np.exp(u)
            out["logj"] += u
        if self.clock != "strict":
            u = Z[:, self.slices["substrates"]]
            out["substrates"] = np.exp(u)
            out["logj"] += u.sum(axis=1)
        if self.coalescent != "constant":
            out["thetas"] = Z[:, self.slices["thetas"]]                 # vector[G] thetas: unconstrained (log space)
        self._constrain_common(Z, out)
        # heights = transform(props, height, map, lowers) and the log-det-Jacobian loop
        # (generate_script.py:711-752), in the library's host code: O(B S), no Python loop over nodes
        from .likelihood import ratios_forward
        with np.errstate(invalid="ignore"):
            ok = np.isfinite(out["props"]).all(axis=1) & np.isfinite(out["height"])
        out["heights"], out["logjac_heights"] = ratios_forward(
            self.map32, self.lowers_or_none, np.where(ok[:, None], out["props"], 0.5),
            np.where(ok, out["height"], self.lower_root + 1.0))
        out["heights"][~ok] = np.nan
        return out

    def constrained_matrix(self, Z: np.ndarray) -> np.ndarray:
        c = self.constrain(Z)
        cols = [c["wshape"][:, None]] if self.C > 1 else []
        cols.append(c["props"])
        cols += [c["rate"][:, None]] if self.clock == "strict" else \
            [c["substrates"]] + [c[k][:, None] for k in self._scalar_names()[0]]
        cols.append(c["height"][:, None])
        cols += [c["theta"][:, None]] if self.coalescent == "constant" else [c["thetas"], c["tau"][:, None]]
        for k in ("rates", "kappa", "freqs"):
            if k in c:
                cols.append(c[k] if c[k].ndim == 2 else c[k][:, None])
        cols.append(c["heights"])
        return np.concatenate(cols, axis=1)

    # span of every branch in time units: heights[parent] - (heights[node] | lowers[node])
    def _spans(self, h):
        return h[:, self.parent_h] - np.where(self.internal[None, :], h[:, self.node_h], self.lowers[self.node][None, :])

    def _events(self, h):
        B, S = h.shape[0], self.S
        times = np.concatenate([np.broadcast_to(self.lowers[:S], (B, S)), h], axis=1)     # indexed by node
        return times

    def _constant_coalescent(self, h, theta, want_grad):
        """constant_coalescent_log (generate_script.py:285-349), batched; gradient w.r.t. heights, theta."""
        B, S = h.shape[0], self.S
        times = self._events(h)
        order = np.argsort(times, axis=1, kind="stable")
        ts = np.take_along_axis(times, order, axis=1)
        delta = np.where(order < S, 1.0, -1.0)                   # sampling event +1, coalescent event -1
        k_before = np.cumsum(delta, axis=1) - delta
        c = 0.5 * k_before * (k_before - 1.0)
        interval = np.diff(ts, axis=1, prepend=ts[:, :1])
        tot = (interval * c).sum(axis=1)
        logp = -tot / theta - (S - 1) * np.log(theta)
        if not want_grad:
            return logp, None, None
        # event i ends the interval weighted by c_i and starts the one weighted by c_{i+1}
        gt_sorted = (-c + np.concatenate([c[:, 1:], np.zeros((B, 1))], axis=1)) / theta[:, None]
        gt = np.empty_like(gt_sorted)
        np.put_along_axis(gt, order, gt_sorted, axis=1)
        return logp, gt[:, S:], tot / theta ** 2 - (S - 1) / theta

    def _skyride_coalescent(self, h, thetas, want_grad):
        """skyride_coalescent_log (generate_script.py:352-414), batched: the interval that ends at the j-th
        coalescent event (and the sampling intervals before it) has population size exp(thetas[j])."""
        B, S = h.shape[0], self.S
        times = self._events(h)
        order = np.argsort(times, axis=1, kind="stable")
        ts = np.take_along_axis(times, order, axis=1)
        coal = order >= S
        delta = np.where(coal, -1.0, 1.0)
        k_before = np.cumsum(delta, axis=1) - delta
        c = 0.5 * k_before * (k_before - 1.0)
        idx = np.minimum(np.cumsum(coal, axis=1) - coal, S - 2)          # coalescent events strictly before the event
        interval = np.diff(ts, axis=1, prepend=ts[:, :1])
        w = c * np.exp(-np.take_along_axis(thetas, idx, axis=1))
        logp = -(interval * w).sum(axis=1) - (np.take_along_axis(thetas, idx, axis=1) * coal).sum(axis=1)
        if not want_grad:
            return logp, None, None
        gt_sorted = -w + np.concatenate([w[:, 1:], np.zeros((B, 1))], axis=1)
        gt = np.empty_like(gt_sorted)
        np.put_along_axis(gt, order, gt_sorted, axis=1)
        rows = np.repeat(np.arange(B), idx.shape[1])
        gth = np.bincount(rows * (S - 1) + idx.ravel(), weights=(interval * w - coal).ravel(),
                          minlength=B * (S - 1)).reshape(B, S - 1)
        return logp, gt[:, S:], gth

    def _skygrid_coalescent(self, h, thetas, want_grad):
        """skygrid_coalescent_log (generate_script.py:440-520), batched: the time axis is cut at the sampling
        / coalescent events AND at the grid points grid[0..G-2]; a segment starting at s has lineage count
        k(s) and population size exp(thetas[#grid points <= s])."""
        B, S, G = h.shape[0], self.S, self.G
        nn = self.nn
        times = self._events(h)
        cuts = np.broadcast_to(self.grid[:G - 1], (B, G - 1))
        allt = np.concatenate([times, cuts], axis=1)
        kind = np.concatenate([np.where(np.arange(nn) < S, 1.0, -1.0), np.zeros(G - 1)])      # +1 tip, -1 coalescence, 0 grid
        # ties: a grid point sorts AFTER an event at the same time (Stan switches only when finish > grid[index])
        order = np.lexsort((np.broadcast_to(kind == 0, allt.shape), allt), axis=1)
        ts = np.take_along_axis(allt, order, axis=1)
        kd = kind[order]
        k_after = np.cumsum(kd, axis=1)                           # lineages after the point
        g_after = np.cumsum(kd == 0, axis=1)                      # grid index (0-based) after the point
        seg = np.diff(ts, axis=1)                                 # segment i: from point i to point i+1
        ka, ga = k_after[:, :-1], g_after[:, :-1]
        c = 0.5 * ka * (ka - 1.0)
        inv_pop = np.exp(-np.take_along_axis(thetas, ga, axis=1))
        w = c * inv_pop                                           # rate of the segment
        coal = kd == -1
        g_at = np.concatenate([np.zeros((B, 1), dtype=np.int64), g_after[:, :-1]], axis=1)   # grid index before the point
        logp = -(seg * w).sum(axis=1) - (np.take_along_axis(thetas, g_at, axis=1) * coal).sum(axis=1)
        if not want_grad:
            return logp, None, None
        # d/d(time of point i) = -w_{i-1} + w_i  (end of segment i-1, start of segment i)
        zero = np.zeros((B, 1))
        gpt = -np.concatenate([zero, w], axis=1) + np.concatenate([w, zero], axis=1)
        gt = np.zeros_like(allt)
        np.put_along_axis(gt, order, gpt, axis=1)
        gth = np.zeros((B, G))
        contrib = seg * w                                          # d/dtheta_g of -(seg c exp(-theta_g)) = +seg w
        for b in range(B) if B <= 4 else ():                      # tiny batches: bincount per row is cheapest
            gth[b] = np.bincount(ga[b], weights=contrib[b], minlength=G) - np.bincount(g_at[b], weights=coal[b], minlength=G)
        if B > 4:
            rows = np.repeat(np.arange(B), ga.shape[1])
            gth = np.bincount(rows * G + ga.ravel(), weights=contrib.ravel(), minlength=B * G).reshape(B, G)
            rows = np.repeat(np.arange(B), g_at.shape[1])
            gth -= np.bincount(rows * G + g_at.ravel(), weights=coal.ravel().astype(float), minlength=B * G).reshape(B, G)
        return logp, gt[:, S:nn], gth

    def _device_front_end(self) -> bool:
        """The handle runs heights -> blens and the reverse sweeps on the device (PHYLO_B200_HOST_FRONT_END=1: keep them here)."""
        import os
        return hasattr(self.lik, "value_grad_ratios_batch") and os.environ.get("PHYLO_B200_HOST_FRONT_END", "0") in ("", "0")

    def _likelihood_ratios(self, sub, rs_, ps_, hbar_extra, want_grad):
        """One ``phylo_b200_eval_ratios_batch`` call: (logL, site-model gradients, d/dprops, d/droot height, d/drate(s)),
        the last three with the tree prior's ``hbar_extra`` and the log-Jacobian included.  A rejected batch is retried
        draw by draw; a rejected draw is -inf with a zero gradient (as ``_likelihood``)."""
        from .likelihood import ValueGrad
        n = sub["props"].shape[0]
        rates = sub["rate"][:, None] if self.clock == "strict" else sub["substrates"]
        subst = self._subst_arg(sub)

        def call(sl):
            cut = lambda a: None if a is None else a[sl]
            return self.lik.value_grad_ratios_batch(self.map32, sub["props"][sl], sub["height"][sl], rates[sl], self.lowers_or_none,
                                                    cut(subst), cut(sub.get("freqs")), rs_[sl], ps_[sl],
                                                    hbar_extra=cut(hbar_extra), want_grad=want_grad, return_heights=False)
        try:
            parts = [call(slice(None))]
        except Exception as e:
            if type(e).__name__ != "PhyloDomainError":
                raise
            parts = []
            for i in range(n):
                try:
                    parts.append(call(slice(i, i + 1)))
                except Exception as e1:
                    if type(e1).__name__ != "PhyloDomainError":
                        raise
                    z = lambda k: np.zeros((1, k))
                    parts.append({"logp": np.full(1, -np.inf), "g_props": z(self.S - 2), "g_root": np.zeros(1),
                                  "g_rates": z(rates.shape[1]),
                                  "rest": ValueGrad(np.full(1, -np.inf), None, z(max(self.lik.nsubst, 0)), z(4), z(self.C), z(self.C))})
        cat = lambda f: np.concatenate([f(p) for p in parts])
        ll = cat(lambda p: p["logp"])
        if not want_grad:
            return ll, None, None, None, None
        vg = ValueGrad(ll, None, cat(lambda p: p["rest"].grad_subst), cat(lambda p: p["rest"].grad_freqs),
                       cat(lambda p: p["rest"].grad_rs), cat(lambda p: p["rest"].grad_ps))
        return ll, vg, cat(lambda p: p["g_props"]), cat(lambda p: p["g_root"]), cat(lambda p: p["g_rates"])

    def log_prob_grad(self, Z: np.ndarray, want_grad: bool = True) -> Tuple[np.ndarray, Optional[np.ndarray]]:
        Z = np.atleast_2d(np.asarray(Z, dtype=np.float64))
        B = Z.shape[0]
        strict, constant = self.clock == "strict", self.coalescent == "constant"
        with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
            c = self.constrain(Z)
            rs, drs, ps = self._site_model(c, B)
            span = self._spans(c["heights"])
            rate_b = c["rate"][:, None] if strict else c["substrates"]
            ok = self._ok_common(c, rs) & np.all(np.isfinite(c["heights"]), axis=1) & np.all(span > 0, axis=1) \
                & np.all(np.isfinite(rate_b) & (rate_b > 0), axis=1) & np.all(rate_b * span < 1e6, axis=1)
            clock_names, coal_names = self._scalar_names()
            for name in clock_names + coal_names:
                ok &= np.isfinite(c[name]) & (c[name] > 0)
            if not constant:
                ok &= np.all(np.abs(c["thetas"]) < 700.0, axis=1)
        lp = np.full(B, -np.inf)
        G = np.zeros((B, self.dim)) if want_grad else None
        if not ok.any():
            return lp, G
        idx = slice(None) if ok.all() else np.nonzero(ok)[0]                     # the usual case: no copy
        sub = c if isinstance(idx, slice) else self._take(c, idx)
        rs_, drs_, ps_, span = rs[idx], drs[idx], ps[idx], span[idx]
        n = span.shape[0]
        h = sub["heights"]
        rate_b = sub["rate"][:, None] if strict else sub["substrates"][:, self.node]     # per pre-order row
        # heights -> blens, the likelihood and the reverse sweeps down to d/dprops run in ONE library call on the
        # device when the handle offers it (phylo_b200_eval_ratios_batch); the tree prior is evaluated first so
        # that its adjoint of the heights rides through the same sweep.  Otherwise (pattern shards over
        # torch.distributed, the CPU tests' stand-in likelihoods) the chain rule is done here.
        device_front = self._device_front_end()
        if not device_front:
            blens = np.empty((n, self.bcount))
            blens[:, self.node] = rate_b * span                                      # generate_script.py:660-679
            args = (blens, self._subst_arg(sub), sub.get("freqs"), rs_, ps_)
            ll, vg = self._likelihood(args, want_grad)
        prior = self._prior_common(sub)
        if constant:
            theta = sub["theta"]
            coal, g_coal_h, g_coal_theta = self._constant_coalescent(h, theta, want_grad)
            prior += -np.log(theta) + coal                                           # theta ~ oneOnX()
        else:
            thetas, tau = sub["thetas"], sub["tau"]
            coal, g_coal_h, g_coal_thetas = (self._skygrid_coalescent if self.coalescent == "skygrid"
                                             else self._skyride_coalescent)(h, thetas, want_grad)
            dth = np.diff(thetas, axis=1)
            ssq = (dth ** 2).sum(axis=1)
            prior += coal + np.log(tau) * (self.G - 1.0) / 2.0 - ssq * tau / 2.0 \
                - (self.G - 1.0) / 2.0 * math.log(2.0 * math.pi)                     # thetas ~ gmrf(tau): user function, constant kept
            prior += (0.001 - 1.0) * np.log(tau) - 0.001 * tau                      # tau ~ gamma(0.001, 0.001)
        if strict:
            prior += -1000.0 * sub["rate"]                                           # rate ~ exponential(1000)
        elif self.clock == "uced":
            s_, mean = sub["substrates"], sub["uced_mean"]
            prior += -self.bcount * np.log(mean) - s_.sum(axis=1) / mean             # substrates ~ exponential(1/uced_mean)
            prior += -1000.0 * mean                                                  # uced_mean ~ exponential(1000)
        else:
            s_, mean, sd = sub["substrates"], sub["ucln_mean"], sub["ucln_stdev"]
            mu = np.log(mean) - 0.5 * sd ** 2
            ls = np.log(s_)
            resid = ls - mu[:, None]
            prior += (-ls - resid ** 2 / (2.0 * sd[:, None] ** 2)).sum(axis=1) - self.bcount * np.log(sd)
            prior += -1000.0 * mean + (0.5396 - 1.0) * np.log(sd) - 2.6184 * sd
        if device_front:
            ll, vg, gp, g_root, g_lik_rates = self._likelihood_ratios(sub, rs_, ps_, g_coal_h if want_grad else None, want_grad)
        lp[idx] = ll + prior + sub["logjac_heights"] + sub["logj"]
        if not want_grad:
            return lp, None
        g = np.zeros((n, self.dim))
        if device_front:
            gb_span = g_lik_rates                                                    # d logL / d rate(s): [n, 1] or node order
        else:
            gb = np.reshape(vg.grad_blens, (n, self.bcount))[:, self.node]           # per pre-order row
            w = rate_b * gb
            hbar = g_coal_h + (w @ self.scatter_blens if isinstance(self.scatter_blens, np.ndarray)
                               else (self.scatter_blens @ w.T).T)
            # reverse sweep of the ratio transform and of its log-Jacobian (library host code)
            from .likelihood import ratios_reverse
            gp, g_root = ratios_reverse(self.map32, self.lowers_or_none, sub["props"], h, hbar)
            if strict:
                gb_span = (gb * span).sum(axis=1)[:, None]
            else:
                gb_span = np.zeros((n, self.bcount))
                gb_span[:, self.node] = gb * span                                    # likelihood, node order
        p = sub["props"]
        g[:, self.slices["props"]] = gp * p * (1.0 - p) + (1.0 - 2.0 * p)
        g[:, self.slices["height"]] = (g_root * (sub["height"] - self.lower_root) + 1.0)[:, None]
        if strict:
            g_rate = gb_span[:, 0] - 1000.0
            g[:, self.slices["rate"]] = (g_rate * sub["rate"] + 1.0)[:, None]
        elif self.clock == "uced":
            gs = gb_span.copy()
            gs += -1.0 / mean[:, None]
            g[:, self.slices["substrates"]] = gs * s_ + 1.0
            g_mean = -self.bcount / mean + s_.sum(axis=1) / mean ** 2 - 1000.0
            g[:, self.slices["uced_mean"]] = (g_mean * mean + 1.0)[:, None]
        else:
            gs = gb_span.copy()
            gs += -1.0 / s_ - resid / (sd[:, None] ** 2 * s_)
            g[:, self.slices["substrates"]] = gs * s_ + 1.0
            r1 = resid.sum(axis=1) / sd ** 2                                         # d prior / d mu
            g_mean = r1 / mean - 1000.0
            g_sd = -r1 * sd + (resid ** 2).sum(axis=1) / sd ** 3 - self.bcount / sd + (0.5396 - 1.0) / sd - 2.6184
            g[:, self.slices["ucln_mean"]] = (g_mean * mean + 1.0)[:, None]
            g[:, self.slices["ucln_stdev"]] = (g_sd * sd + 1.0)[:, None]
        if constant:
            g[:, self.slices["theta"]] = ((g_coal_theta - 1.0 / theta) * theta + 1.0)[:, None]
        else:
            gth = g_coal_thetas.copy()
            gth[:, :-1] += tau[:, None] * dth
            gth[:, 1:] -= tau[:, None] * dth
            g[:, self.slices["thetas"]] = gth
            g_tau = (self.G - 1.0) / (2.0 * tau) - ssq / 2.0 + (0.001 - 1.0) / tau - 0.001
            g[:, self.slices["tau"]] = (g_tau * tau + 1.0)[:, None]
        self._grad_common(g, sub, vg, drs_)
        G[idx] = g
        return lp, G


class StrictClockModel(ClockModel):
    """``ClockModel`` with a strict clock and a constant-size coalescent -- the fluA quick start
    (BASELINE config 1; tests/golden/fluA-HKY-W4-external.stan)."""

    def __init__(self, lik, model: str, map_, lowers=None, lower_root: Optional[float] = None, rates_alpha=None,
                 freqs_alpha=None):
        super().__init__(lik, model, map_, lowers, lower_root, rates_alpha, freqs_alpha, "strict", "constant")


# ---------------------------------------------------------------------------------------------------
# mean-field and full-rank ADVI
# ---------------------------------------------------------------------------------------------------
@dataclass
class MeanFieldFit:
    """Result of ``advi`` (either family; ``omega`` is None for full rank, ``L`` None for mean field)."""
    mu: np.ndarray
    omega: Optional[np.ndarray]
    eta: float
    iterations: int
    converged: bool
    elbo_trace: List[Tuple[int, float]] = field(default_factory=list)
    draws: Optional[np.ndarray] = None          # [output_samples, n constrained], columns = names
    names: List[str] = field(default_factory=list)
    likelihood_calls: int = 0                    # library calls (each one a whole batch of draws)
    likelihood_draws: int = 0                    # draws evaluated in total
    L: Optional[np.ndarray] = None               # Cholesky factor of the covariance (full rank)

    def mean(self) -> Dict[str, float]:
        return dict(zip(self.names, self.draws.mean(axis=0))) if self.draws is not None else {}


class _MeanFieldFamily:
    """zeta = mu + exp(omega) * eta   (stan::variational::normal_meanfield); theta = [mu | omega]."""

    def __init__(self, dim):
        self.d = dim

    def init(self, mu):
        return np.concatenate([mu, np.zeros(self.d)])

    def transform(self, th, eta):
        return th[:self.d] + np.exp(th[self.d:]) * eta

    def entropy(self, th):
        return 0.5 * self.d * (1.0 + math.log(2.0 * math.pi)) + th[self.d:].sum()

    def grad(self, th, g, eta):
        return np.concatenate([g.mean(axis=0), (g * eta).mean(axis=0) * np.exp(th[self.d:]) + 1.0])


class _FullRankFamily:
    """zeta = mu + L eta, L lower triangular, initialised to I  (stan::variational::normal_fullrank);
    theta = [mu | L row-major]; entries above the diagonal have zero gradient and stay zero."""

    def __init__(self, dim):
        self.d = dim

    def init(self, mu):
        return np.concatenate([mu, np.eye(self.d).ravel()])

    def _L(self, th):
        return th[self.d:].reshape(self.d, self.d)

    def transform(self, th, eta):
        return th[:self.d] + eta @ self._L(th).T

    def entropy(self, th):
        return 0.5 * self.d * (1.0 + math.log(2.0 * math.pi)) + np.log(np.abs(np.diag(self._L(th)))).sum()

    def grad(self, th, g, eta):
        gl = np.tril(g.T @ eta / g.shape[0])
        gl[np.diag_indices(self.d)] += 1.0 / np.diag(self._L(th))
        return np.concatenate([g.mean(axis=0), gl.ravel()])


class _Advi:
    ALPHA, TAU = 0.1, 1.0
    ETA_SEQUENCE = (100.0, 10.0, 1.0, 0.1, 0.01)

    def __init__(self, model, family, grad_samples, elbo_samples, rng):
        self.m, self.f, self.ng, self.ne, self.rng = model, family, int(grad_samples), int(elbo_samples), rng
        self.calls = self.draws = 0

    def elbo(self, th) -> float:
        eta = self.rng.standard_normal((self.ne, self.m.dim))
        lp = self.m.log_prob(self.f.transform(th, eta))
        self.calls += 1
        self.draws += self.ne
        lp = lp[np.isfinite(lp)]                                         # Stan drops failed draws
        if lp.size == 0:
            return -np.inf
        return float(lp.mean() + self.f.entropy(th))

    def elbo_grad(self, th):
        eta = self.rng.standard_normal((self.ng, self.m.dim))
        lp, g = self.m.log_prob_grad(self.f.transform(th, eta))
        self.calls += 1
        self.draws += self.ng
        if not np.all(np.isfinite(lp)) or not np.all(np.isfinite(g)):
            raise FloatingPointError("non-finite log density or gradient in an ELBO gradient draw")
        return self.f.grad(th, g, eta)

    def ascend(self, th, eta, iters, on_check=None, eval_elbo=0):
        hist = None
        for it in range(1, iters + 1):
            g = self.elbo_grad(th)
            hist = g * g if hist is None else self.ALPHA * g * g + (1.0 - self.ALPHA) * hist
            th = th + eta / math.sqrt(it) / (self.TAU + np.sqrt(hist)) * g
            if eval_elbo and it % eval_elbo == 0 and on_check(it, th):
                return th, it, True
        return th, iters, False

    def adapt_eta(self, th0, adapt_iter):
        elbo_init = self.elbo(th0)
        best, eta_best = -np.inf, None
        for k, eta in enumerate(self.ETA_SEQUENCE):
            try:
                th, _, _ = self.ascend(th0.copy(), eta, adapt_iter)
                e = self.elbo(th)
            except FloatingPointError:
                e = -np.inf
            if e < best and best > elbo_init:
                break
            if e > best:
                best, eta_best = e, eta
            if k == len(self.ETA_SEQUENCE) - 1 and not best > elbo_init:
                raise RuntimeError("all proposed step sizes failed; the model may be ill-conditioned")
        return eta_best


def advi(model, *, algorithm: str = "meanfield", iter: int = 10000, grad_samples: int = 1, elbo_samples: int = 100,
         eval_elbo: int = 100, tol_rel_obj: float = 0.001, eta: Optional[float] = None, adapt_iter: int = 50,
         output_samples: int = 1000, seed: int = 1, init="random", verbose: bool = False) -> MeanFieldFit:
    """ADVI with a mean-field or full-rank Gaussian family; keyword names follow ``pystan.StanModel.vb``
    (phylostan/phylostan.py:311-313; ``algorithm`` is phylostan's ``-q/--variational``).

    ``init``: "random" (Stan's uniform(-2, 2) on the unconstrained scale), "zero", or an unconstrained
    vector.  ``grad_samples`` draws per iteration and ``elbo_samples`` draws per ELBO estimate are each
    evaluated by one batched library call."""
    if algorithm not in ("meanfield", "fullrank"):
        raise ValueError("algorithm must be meanfield or fullrank")
    rng = np.random.default_rng(seed)
    if isinstance(init, str):
        mu = rng.uniform(-2.0, 2.0, model.dim) if init == "random" else np.zeros(model.dim)
    else:
        mu = np.asarray(init, dtype=np.float64).copy()
        if mu.shape != (model.dim,):
            raise ValueError(f"init must have {model.dim} entries")
    fam = _MeanFieldFamily(model.dim) if algorithm == "meanfield" else _FullRankFamily(model.dim)
    th = fam.init(mu)
    A = _Advi(model, fam, grad_samples, elbo_samples, rng)
    if eta is None:
        eta = A.adapt_eta(th, adapt_iter)
    trace: List[Tuple[int, float]] = []
    cb_size = max(int(0.1 * iter / eval_elbo), 2)
    ring: List[float] = []
    state = {"prev": A.elbo(th)}
    trace.append((0, state["prev"]))

    def on_check(it, th_):
        e = A.elbo(th_)
        trace.append((it, e))
        delta = abs((e - state["prev"]) / e) if e != 0 else np.inf
        state["prev"] = e
        ring.append(delta)
        del ring[:-cb_size]
        if verbose:
            print(f"  {it:6d}  ELBO {e:.3f}  delta_mean {np.mean(ring):.5f}  delta_med {np.median(ring):.5f}")
        return np.mean(ring) < tol_rel_obj or np.median(ring) < tol_rel_obj

    th, its, conv = A.ascend(th, eta, iter, on_check=on_check, eval_elbo=eval_elbo)
    d = model.dim
    fit = MeanFieldFit(th[:d].copy(), th[d:].copy() if algorithm == "meanfield" else None, float(eta), its, conv, trace,
                       names=model.constrained_names(), likelihood_calls=A.calls, likelihood_draws=A.draws,
                       L=th[d:].reshape(d, d).copy() if algorithm == "fullrank" else None)
    if output_samples:
        fit.draws = model.constrained_matrix(fam.transform(th, rng.standard_normal((output_samples, d))))
    return fit


def advi_meanfield(model, **kw) -> MeanFieldFit:
    """``advi(model, algorithm="meanfield", ...)``."""
    return advi(model, algorithm="meanfield", **kw)


# ---------------------------------------------------------------------------------------------------
# command line: tree + alignment in, posterior-draw CSV out
# ---------------------------------------------------------------------------------------------------
def main(argv=None) -> int:
    import argparse
    import csv

    from . import encode, likelihood

    ap = argparse.ArgumentParser(description="batched ADVI / NUTS on one B200: unrooted tree, or time tree with a strict or "
                                             "uncorrelated-lognormal clock and a constant or skygrid coalescent "
                                             "(option names of `phylostan run`)")
    ap.add_argument("-t", "--tree", required=True)
    ap.add_argument("-i", "--input", required=True, help="alignment (FASTA or NEXUS)")
    ap.add_argument("-m", "--model", default="GTR", choices=("JC69", "HKY", "GTR"))
    ap.add_argument("-C", "--categories", type=int, default=1)
    ap.add_argument("--clock", choices=("strict", "ucln", "uced"), help="time tree with this clock (omit: unrooted tree)")
    ap.add_argument("-c", "--coalescent", default="constant", choices=("constant", "skyride", "skygrid"))
    ap.add_argument("--grid", type=int, help="number of grid points in skygrid")
    ap.add_argument("--cutoff", type=float, help="a cutoff for skygrid")
    ap.add_argument("--heterochronous", action="store_true", help="tip dates from the tree's root-to-tip distances")
    ap.add_argument("--dates", help="csv file with header name,date")
    ap.add_argument("-o", "--output", required=True, help="CSV of draws from the approximation")
    ap.add_argument("--iter", type=int, default=10000)
    ap.add_argument("--grad_samples", type=int, default=1)
    ap.add_argument("--elbo_samples", type=int, default=100)
    ap.add_argument("--tol_rel_obj", type=float, default=0.001)
    ap.add_argument("-e", "--eta", type=float)
    ap.add_argument("-a", "--algorithm", default="vb", choices=("vb", "nuts", "hmc"))
    ap.add_argument("--chains", type=int, default=8, help="chains of -a hmc, advanced in lock step (one batched "
                                                              "library call per leapfrog step)")
    ap.add_argument("-q", "--variational", default="meanfield", choices=("meanfield", "fullrank"))
    ap.add_argument("--samples", type=int, default=1000)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args(argv)
    tree, seqs = encode.read_tree(a.tree), encode.read_alignment(a.input)
    rooted = a.clock is not None
    enc = encode.encode(tree, seqs, rooted=rooted)
    with likelihood.TreeLikelihood(enc.peel, enc.tipmask, enc.weights, model=a.model, categories=a.categories,
                                   rooted=rooted) as lik:
        if rooted:
            dates = None
            if a.dates:
                with open(a.dates) as f:
                    dates = {row["name"]: float(row["date"].strip()) for row in csv.DictReader(f)}
            oldest = encode.setup_dates(tree, dates, a.heterochronous)
            lowers = encode.get_lowers(tree) if oldest is not None else None
            grid = np.linspace(0, a.cutoff, a.grid)[1:] if a.coalescent == "skygrid" else None   # phylostan.py:275-277
            model = ClockModel(lik, a.model, enc.map, lowers, clock=a.clock, coalescent=a.coalescent, grid=grid)
        else:
            model = UnrootedModel(lik, a.model)
        if a.algorithm == "nuts":           # pystan's convention: iter = warm-up + sampling, half each
            from .sampling import nuts
            fit = nuts(model, num_warmup=a.iter // 2, num_samples=a.iter - a.iter // 2, seed=a.seed, verbose=True)
            print(f"step size {fit.stepsize:.4g}  mean tree depth {fit.treedepth.mean():.2f}  "
                  f"divergent {int(fit.divergent.sum())}  gradient evaluations {fit.gradient_evaluations}")
        elif a.algorithm == "hmc":
            from .sampling import hmc
            fit = hmc(model, chains=a.chains, num_warmup=a.iter // 2, num_samples=a.iter - a.iter // 2, seed=a.seed,
                      verbose=True)
            print(f"step size {fit.stepsize:.4g}  leapfrog steps {fit.n_leapfrog}  accept {fit.accept_stat.mean():.2f}  "
                  f"batched library calls {fit.gradient_calls}")
            fit.draws = fit.draws.reshape(-1, fit.draws.shape[-1])
        else:
            fit = advi(model, algorithm=a.variational, iter=a.iter, grad_samples=a.grad_samples,
                       elbo_samples=a.elbo_samples, tol_rel_obj=a.tol_rel_obj, eta=a.eta, output_samples=a.samples,
                       seed=a.seed, verbose=True)
            print(f"eta {fit.eta}  iterations {fit.iterations}  converged {fit.converged}  "
                  f"library calls {fit.likelihood_calls} ({fit.likelihood_draws} draws)")
    with open(a.output, "w") as f:
        f.write(",".join(fit.names) + "\n")
        for row in fit.draws:
            f.write(",".join(f"{v:.9g}" for v in row) + "\n")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
