#!/usr/bin/env python3
"""Golden Stan programs for the external-likelihood switch (authoring container only): runs the
REFERENCE's generator /root/reference/phylostan/generate_script.py::get_model unmodified and the repo's
phylostan_b200.generate.externalize on its output.  Writes tests/golden/*-external.stan.

    python tests/golden/make_golden_stan.py
"""
import argparse
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
from phylostan_b200 import generate as G  # noqa: E402


def load_reference():
    spec = importlib.util.spec_from_file_location("ref_generate_script", "/root/reference/phylostan/generate_script.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def params(**kw):
    d = dict(model="GTR", invariant=False, categories=4, heterogeneity="weibull", heterochronous=True, clock="strict",
             estimate_rate=True, coalescent="constant", speciation=None, grid=None, cutoff=None, geo=False,
             rescaling_geo=False)
    d.update(kw)
    return argparse.Namespace(**d)


CASES = {
    # README quick start: phylostan build -s fluA.stan -m HKY -C 4 --heterochronous --estimate_rate --clock strict -c constant
    "fluA-HKY-W4": params(model="HKY"),
    # DS1 paper run: unrooted GTR + W4 (examples/SConstruct:158-203)
    "DS1-GTR-W4": params(clock=None, coalescent=None, heterochronous=False, estimate_rate=False),
}

if __name__ == "__main__":
    ref = load_reference()
    for name, p in CASES.items():
        out = os.path.join(HERE, name + "-external.stan")
        with open(out, "w") as fp:
            fp.write(G.get_model(p, ref))
        print("wrote", out)
