"""Run-side glue for the reference's command line (phylostan/phylostan.py): what `phylostan build` / `phylostan run`
call when ``--gpu-likelihood`` is given.  The reference file only needs the hunks of
``integration/phylostan-gpu-likelihood.patch``; everything they call lives here.

    phylostan build -s m.stan -m GTR -C 4 --gpu-likelihood               # generate_script.get_model + the external call
    phylostan run   -s m.stan ... --gpu-likelihood [--gpu-devices 0,1]   # compile with the shim, publish the data, run

* ``add_arguments``  flags for ``create_build_parser`` (phylostan.py:69-95; the run parser is built from it, :31)
* ``build_script``   replaces ``get_model(arg)`` in ``build`` (:149-152)
* ``stan_model``     replaces ``pystan.StanModel(file=arg.script)`` (:158, :296): dlopen(libphylo_b200, RTLD_GLOBAL)
                     first, then compile with ``allow_undefined`` + the shim header (the eigen/eigen.py:79-87 mechanism)
* ``publish``        after the Stan data dict is complete (:183-287), before ``sm.vb`` / ``sm.sampling`` (:311, :319):
                     creates the device handle from that dict -- on several GPUs with ``--gpu-devices`` -- and publishes it

pystan is imported only inside ``stan_model``; the module itself needs neither pystan nor a GPU.
"""
from __future__ import annotations

import os
from typing import Any, Dict, List, Optional

import numpy as np

from . import generate


def add_arguments(parser) -> None:
    parser.add_argument("--gpu-likelihood", dest="gpu_likelihood", action="store_true",
                        help="evaluate the tree likelihood and its gradient with libphylo_b200 (external Stan function)")
    parser.add_argument("--gpu-devices", dest="gpu_devices", default="0",
                        help="comma-separated CUDA ordinals; more than one shards the site patterns over those GPUs "
                             "[default: %(default)s]")
    parser.add_argument("--gpu-heights", dest="gpu_heights", action="store_true",
                        help="clock trees: also move the heights -> branch-length loop into the library")


def enabled(arg) -> bool:
    return bool(getattr(arg, "gpu_likelihood", False))


def devices(arg) -> List[int]:
    txt = str(getattr(arg, "gpu_devices", "0") or "0")
    out = [int(x) for x in txt.split(",") if x.strip() != ""]
    if not out or min(out) < 0:
        raise ValueError("--gpu-devices expects a comma-separated list of CUDA ordinals")
    return out


def build_script(arg, generate_script=None) -> str:
    """The Stan program of ``phylostan build``: the reference's own generator output with the P-matrix functions and
    the pruning loops swapped for one external call (phylostan_b200.generate)."""
    return generate.get_model(arg, generate_script, heights=bool(getattr(arg, "gpu_heights", False)) and arg.clock is not None)


def preload() -> None:
    """dlopen(libphylo_b200.so, RTLD_GLOBAL): must happen before a compiled (or un-pickled, phylostan.py:301) model
    extension is imported, so that its undefined ``phylo_b200_*`` symbols bind to this copy."""
    from . import likelihood
    likelihood.lib()


def stan_model(script_path: str, pystan=None):
    """``pystan.StanModel(file=...)`` with the shim compiled in."""
    from . import likelihood
    likelihood.lib()  # RTLD_GLOBAL, before the model extension is imported: its phylo_b200_* symbols bind to this copy
    if pystan is None:
        import pystan  # noqa: PLC0415  (absent in the build image; present where phylostan runs)
    return pystan.StanModel(file=script_path, **generate.stan_model_kwargs())


def publish(arg, data: Dict[str, Any]):
    """Device handle from the Stan data dict + ``phylo_b200_set_default``.  Keep the returned object alive while Stan runs.
    ``--gpu-devices a,b,...`` -> one handle over those GPUs (phylo_b200_create_multi)."""
    from . import likelihood as lk
    devs = devices(arg)
    tipdata = np.asarray(data["tipdata"], dtype=np.float64)
    if tipdata.ndim != 3 or tipdata.shape[2] != 4:
        raise ValueError("data['tipdata'] must be [S, L, 4] (phylostan/phylostan.py:183-204)")
    # the reference's one-hot / all-ones rows -> 4-bit state masks (bit s <=> tipdata[.., s] != 0)
    mask = ((tipdata != 0.0) * np.array([1, 2, 4, 8])).sum(axis=2).astype(np.uint8)
    lik = lk.TreeLikelihood(np.asarray(data["peel"]), mask, np.asarray(data["weights"], dtype=np.float64), model=arg.model,
                            categories=int(data.get("C", 1)), rooted=arg.clock is not None,
                            device=devs[0], devices=devs if len(devs) > 1 else None)
    lk.set_default(lik)
    return lik


PATCH = os.path.join(generate.REPO, "integration", "phylostan-gpu-likelihood.patch")


def patched_checkout(dest: str, source: Optional[str] = None) -> str:
    """Copy the importable ``phylostan`` package (or ``source``/phylostan) to ``dest`` and apply the patch there.
    Returns ``dest`` (put it first on sys.path).  Uses patch(1)."""
    import importlib.util
    import shutil
    import subprocess
    if source is None:
        spec = importlib.util.find_spec("phylostan")
        if spec is None or not spec.submodule_search_locations:
            raise ImportError("phylostan is not importable: install it or pass its checkout directory")
        source = os.path.dirname(list(spec.submodule_search_locations)[0])
    shutil.copytree(os.path.join(source, "phylostan"), os.path.join(dest, "phylostan"))
    r = subprocess.run(["patch", "-p1", "-d", dest, "-i", PATCH], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("the patch does not apply to this phylostan:\n" + r.stdout + r.stderr)
    return dest


def main(argv: Optional[List[str]] = None) -> int:
    """``python -m phylostan_b200.run build|run ... --gpu-likelihood``: the reference's command line from a patched
    temporary copy of the installed phylostan package (needs pystan and dendropy, as phylostan itself does)."""
    import importlib
    import sys
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        sys.path.insert(0, patched_checkout(tmp))
        for name in [m for m in sys.modules if m == "phylostan" or m.startswith("phylostan.")]:
            del sys.modules[name]
        ps = importlib.import_module("phylostan.phylostan")
        if argv is not None:
            sys.argv = ["phylostan"] + list(argv)
        ps.main()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
