"""The external-likelihood switch for the reference's Stan generator, and the `phylostan run` plumbing.

The reference emits the pruning recursion as Stan text (``likelihood()``,
phylostan/generate_script.py:961-1055) preceded by ``pmats = calculate_*_p_matrices(...)``
(:1424-1446).  Its experimental ``eigen/`` path instead declares an undefined function and calls it once
(``real pruning_loglik(vector blens);`` eigen/example.stan:3, ``target += pruning_loglik(blens)`` :134)
and compiles with ``allow_undefined=True, includes=[...]`` (eigen/eigen.py:79-87).  ``get_model`` below
applies that mechanism to every model the generator can emit: it calls the reference's own
``get_model(params)`` and swaps the P-matrix + likelihood text for one call of

    real phylo_loglik(vector blens, vector subst, vector freqs, vector rs, vector ps);

implemented by phylostan_b200/stan/phylo_b200_stan.hpp on top of libphylo_b200.so.  Priors, clocks,
coalescent models, the height transform and its Jacobian stay exactly as generated (Stan differentiates
them; the shim returns d/dblens, d/drates|kappa, d/dfreqs, d/drs, d/dps through precomputed_gradients).
"""
from __future__ import annotations

import importlib
import os
from typing import Any, Dict, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(_HERE)

DECLARATION = "\treal phylo_loglik(vector blens, vector subst, vector freqs, vector rs, vector ps);\n"


def _reference_generator(generate_script=None):
    if generate_script is not None:
        return generate_script
    return importlib.import_module("phylostan.generate_script")


def is_mixture(params) -> bool:
    return params.categories > 1 or params.invariant  # generate_script.py:1448


def likelihood_call(params) -> str:
    """The single statement that replaces ``likelihood(mixture, clock)``."""
    subst = {"GTR": "rates", "HKY": "rep_vector(kappa, 1)", "JC69": "rep_vector(0.0, 0)"}[params.model]
    site = "rs, ps" if is_mixture(params) else "rep_vector(1.0, 1), rep_vector(1.0, 1)"
    return "\ttarget += phylo_loglik(blens, {}, freqs, {});\n".format(subst, site)


DECLARATION_HEIGHTS = ("\treal phylo_loglik_heights(real[] heights, real[] rates, int[,] map, real[] lowers, "
                       "vector subst, vector freqs, vector rs, vector ps);\n")
DECLARATION_AUTOCORR = DECLARATION_HEIGHTS.replace("phylo_loglik_heights(", "phylo_loglik_heights_autocorr(")
AUTOCORRELATED = ("acln", "acg", "ace", "aoup", "hsmrf", "gmrf")  # generate_script.py:1333


def heights_call(params) -> str:
    """Replaces ``heights_to_blens`` (generate_script.py:660-679) AND the likelihood: node heights and
    rate(s) go straight to the library, which returns d/dheights and d/drate(s)."""
    subst = {"GTR": "rates", "HKY": "rep_vector(kappa, 1)", "JC69": "rep_vector(0.0, 0)"}[params.model]
    site = "rs, ps" if is_mixture(params) else "rep_vector(1.0, 1), rep_vector(1.0, 1)"
    strict = params.clock == "strict" or not params.estimate_rate       # generate_script.py:1336
    autocorr = params.clock in AUTOCORRELATED
    rate = "substrates" if autocorr or not strict else "rep_array(rate, 1)"
    lowers = "lowers" if params.heterochronous else "rep_array(0.0, 2*S-1)"
    return "\ttarget += phylo_loglik_heights{}(heights, {}, map, {}, {}, freqs, {});\n".format(
        "_autocorr" if autocorr else "", rate, lowers, subst, site)


def externalize(script: str, params, generate_script=None, heights: bool = False) -> str:
    """Rewrite a program produced by the reference's ``get_model(params)``.

    ``heights=True`` (clock trees) also moves the heights -> branch-length loop into the library."""
    g = _reference_generator(generate_script)
    mixture, clock = is_mixture(params), params.clock is not None
    like = g.likelihood(mixture, clock)
    if like not in script:
        raise ValueError("the program does not contain the reference's likelihood block")
    if heights:
        if not clock:
            raise ValueError("heights=True needs a clock tree")
        autocorr = params.clock in AUTOCORRELATED
        strict = params.clock == "strict" or not params.estimate_rate
        h2b = (g.heights_to_blens_autocorr(params.heterochronous) if autocorr
               else g.heights_to_blens(params.heterochronous, strict))
        if h2b not in script:
            raise ValueError("the program does not contain the reference's heights_to_blens block")
        script = script.replace(h2b, "").replace(like, heights_call(params))
        script = script.replace("\tvector [bcount] blens; // branch lengths\n", "")
    else:
        script = script.replace(like, likelihood_call(params))
    # the P-matrix function, its call and the arrays only the recursion used
    fn = {"GTR": g.GTR, "HKY": g.HKY, "JC69": g.JC69}[params.model](params.categories, params.invariant)
    if fn not in script:
        raise ValueError("the program does not contain the reference's P-matrix function")
    script = script.replace(fn, DECLARATION if not heights else
                            DECLARATION_AUTOCORR if params.clock in AUTOCORRELATED else DECLARATION_HEIGHTS)
    kept = []
    for line in script.split("\n"):
        s = line.strip()
        if s.startswith("pmats = calculate_") or s.startswith("vector[4] partials[") or \
                s.startswith("matrix[4,4] pmats[") or s == "real probs[C];":
            continue
        kept.append(line)
    return "\n".join(kept)


def get_model(params, generate_script=None, heights: bool = False) -> str:
    """Drop-in for ``phylostan.generate_script.get_model`` (:1168-1484) with the GPU likelihood."""
    if getattr(params, "geo", False):
        raise ValueError("the phylogeography model (--geo) is outside the accelerated path")
    g = _reference_generator(generate_script)
    return externalize(g.get_model(params), params, g, heights)


def stan_model_kwargs() -> Dict[str, Any]:
    """Keyword arguments for ``pystan.StanModel`` (phylostan/phylostan.py:296), as eigen/eigen.py:81-87.

    pystan 2.19 offers no link-argument hook; libphylo_b200.so is instead loaded RTLD_GLOBAL by
    ``phylostan_b200.likelihood.lib()`` before the compiled model module is imported, so the model's
    undefined ``phylo_b200_*`` symbols bind to it at load time."""
    return {
        "allow_undefined": True,
        "includes": ["phylo_b200_stan.hpp"],
        "include_dirs": [os.path.join(_HERE, "stan"), os.path.join(REPO, "include")],
        "extra_compile_args": ["--std=c++14", "-Wno-int-in-bool-context"],
    }


def publish(data: Dict[str, Any], model: str, clock: Optional[str], device: int = 0):
    """Create the GPU handle from the reference's Stan data dict (phylostan/phylostan.py:183-272) and
    publish it for the shim.  Returns the TreeLikelihood (keep it alive while Stan runs)."""
    from . import likelihood as lk
    tipdata = np.asarray(data["tipdata"])
    lik = lk.TreeLikelihood(np.asarray(data["peel"]), tipdata=tipdata, weights=np.asarray(data["weights"], dtype=float),
                            model=model, categories=int(data.get("C", 1)), rooted=clock is not None, device=device)
    lk.set_default(lik)
    return lik
