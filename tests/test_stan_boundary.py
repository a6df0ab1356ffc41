"""The Stan-facing boundary: generator switch (CPU, against the reference generator when mounted and
against committed golden programs) and the C++ shim (compiled against stub Stan/Eigen headers; run on
the GPU against the oracle)."""
import importlib.util
import json
import os
import struct
import subprocess
import sys

import numpy as np
import pytest

from phylostan_b200 import generate as G

from conftest import GOLDEN, ROOT

REF_GEN = "/root/reference/phylostan/generate_script.py"
sys.path.insert(0, GOLDEN)
from make_golden_stan import CASES, params  # noqa: E402


def _ref():
    if not os.path.exists(REF_GEN):
        pytest.skip("reference tree not mounted (GPU box)")
    spec = importlib.util.spec_from_file_location("ref_generate_script", REF_GEN)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("name", sorted(CASES))
def test_generator_switch_matches_golden(name):
    got = G.get_model(CASES[name], _ref())
    assert got == open(os.path.join(GOLDEN, name + "-external.stan")).read()


def test_golden_programs_call_the_external_function_once():
    for name in CASES:
        s = open(os.path.join(GOLDEN, name + "-external.stan")).read()
        assert s.count("phylo_loglik(") == 2                      # declaration + one call (eigen/example.stan:3,134)
        assert "real phylo_loglik(vector blens, vector subst, vector freqs, vector rs, vector ps);" in s
        assert "target += phylo_loglik(blens," in s
        for gone in ("calculate_", "pmats", "partials[", "probs["):
            assert gone not in s
        # everything upstream of the likelihood stays in Stan
        assert "~ exponential" in s or "~ dirichlet" in s


def test_generator_switch_all_model_variants():
    ref = _ref()
    for kw in (dict(model="HKY"), dict(model="JC69", categories=1), dict(model="JC69", categories=4),
               dict(clock=None, coalescent=None, heterochronous=False, estimate_rate=False),
               dict(invariant=True, categories=1), dict(heterogeneity="discrete"), dict(clock="ucln"),
               dict(coalescent="skygrid", grid=10, cutoff=50.0), dict(coalescent="skyride")):
        p = params(**kw)
        s = G.get_model(p, ref)
        base = ref.get_model(p)
        assert s.count("phylo_loglik(") == 2 and "pmats" not in s
        # only the likelihood changed: priors / transforms / Jacobian lines all survive
        for line in base.split("\n"):
            t = line.strip()
            if t.startswith(("heights ~", "rates ~", "freqs ~", "kappa ~", "wshape ~", "target += log(heights")):
                assert line in s
    with pytest.raises(ValueError):
        G.get_model(params(geo=True), ref)


def test_generator_switch_heights_variant():
    """heights=True also replaces generate_script.py:660-679 / 682-708 (the heights -> blens loops)."""
    ref = _ref()
    for kw in (dict(model="HKY"), dict(clock="ucln"), dict(heterochronous=False, estimate_rate=False)):
        p = params(**kw)
        s = G.get_model(p, ref, heights=True)
        body = s.split("model{")[1]
        assert s.count("phylo_loglik_heights(") == 2 and "blens" not in body and "pmats" not in s
        assert ("lowers" in body) == p.heterochronous or "rep_array(0.0, 2*S-1)" in body
        assert "target += log(heights" in body           # the Jacobian of the height transform stays in Stan
    for clock in ("acln", "ace", "gmrf"):          # generate_script.py:682-708
        p = params(clock=clock)
        s = G.get_model(p, ref, heights=True)
        body = s.split("model{")[1]
        assert s.count("phylo_loglik_heights_autocorr(") == 2 and "phylo_loglik_heights(" not in s
        assert "blens" not in body and "target += phylo_loglik_heights_autocorr(heights, substrates, map," in body
        assert "substrates" in s.split("model{")[0]                # the clock model itself stays in Stan
    with pytest.raises(ValueError):
        G.get_model(params(clock=None, coalescent=None, heterochronous=False, estimate_rate=False), ref, heights=True)


def test_stan_model_kwargs_point_at_existing_files():
    kw = G.stan_model_kwargs()
    assert kw["allow_undefined"] is True
    for inc in kw["includes"]:
        assert any(os.path.exists(os.path.join(d, inc)) for d in kw["include_dirs"])
    assert any(os.path.exists(os.path.join(d, "phylo_b200.h")) for d in kw["include_dirs"])


# ------------------------------------------------------------------------------- C++ shim

def _build_driver(tmp_path):
    exe = str(tmp_path / "shim_driver")
    libdir = os.path.join(ROOT, "phylostan_b200", "csrc")
    cmd = ["g++", "-std=c++14", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "tests", "stubs"),
           "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "phylostan_b200", "stan"),
           os.path.join(ROOT, "tests", "stubs", "shim_driver.cpp"), "-o", exe, "-L" + libdir, "-lphylo_b200",
           "-Wl,-rpath," + libdir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_shim_compiles_as_cxx14_and_fails_loudly_without_handle(tmp_path):
    """pystan compiles models with --std=c++14 (eigen/eigen.py:86); no published handle -> runtime_error."""
    exe = _build_driver(tmp_path)
    r = subprocess.run([exe, "--no-gpu"], capture_output=True, text=True)
    assert r.returncode == 0 and "no default handle" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("name,model,rooted,C,devices", [("fluA", 1, True, 4, ""), ("DS1", 2, False, 4, ""), ("HCV", 0, True, 1, ""),
                                                         ("DS1", 2, False, 4, "multi"), ("fluA", 1, True, 4, "multi")])
def test_shim_value_and_gradient_through_stan_types(tmp_path, datasets, name, model, rooted, C, devices):
    """devices == "multi": the same Stan-facing calls over ONE multi-device handle (phylo_b200_create_multi): every
    visible GPU, or three pattern shards on GPU 0 when the box has one."""
    from oracle import oracle as O
    from phylostan_b200 import encode as E
    d = datasets[name]
    S, L = d["tipmask"].shape
    rng = np.random.default_rng(17)
    bcount = 2 * S - 2 if rooted else 2 * S - 3
    bl = rng.exponential(0.05, bcount) + 1e-4
    subst = {0: np.zeros(0), 1: np.array([4.4]), 2: rng.dirichlet(np.ones(6))}[model]
    fr = rng.dirichlet(np.ones(4) * 5) if model else np.full(4, 0.25)
    rs = E.weibull_rates(0.5, C) if C > 1 else np.ones(1)
    ps = rng.dirichlet(np.ones(C) * 3) if C > 1 else np.ones(1)
    path = tmp_path / "problem.bin"
    with open(path, "wb") as f:
        f.write(struct.pack("5i", S, L, C, model, 1 if rooted else 0))
        f.write(np.ascontiguousarray(d["peel"], dtype=np.int32).tobytes())
        f.write(np.ascontiguousarray(d["tipmask"], dtype=np.uint8).tobytes())
        for a in (d["weights"], bl, subst, fr, rs, ps):
            f.write(np.ascontiguousarray(a, dtype=np.float64).tobytes())
    args = [str(path)]
    if rooted:  # clock-tree front end (phylo_loglik_heights): heights consistent with bl at rate 0.004
        rate = 0.004
        height = {2 * S - 1: 0.0}
        depth = {2 * S - 1: 0.0}
        for node, par in d["map"][1:]:
            depth[int(node)] = depth[int(par)] + bl[int(node) - 1] / rate
        top = max(depth.values())
        hts = np.array([top - depth[S + 1 + k] for k in range(S - 1)])
        lowers = np.array([top - depth[k] if k <= S else 0.0 for k in range(1, 2 * S)])
        hpath = tmp_path / "heights.bin"
        with open(hpath, "wb") as f:
            f.write(np.ascontiguousarray(d["map"], dtype=np.int32).tobytes())
            for a in (hts, lowers, np.array([rate])):
                f.write(np.ascontiguousarray(a, dtype=np.float64).tobytes())
        args.append(str(hpath))
    exe = _build_driver(tmp_path)
    env = dict(os.environ)
    if devices == "multi":
        import torch
        n = torch.cuda.device_count()
        env["PHYLO_SHIM_DEVICES"] = ",".join(str(i) for i in range(n)) if n > 1 else "0,0,0"
    r = subprocess.run([exe] + args, capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    got = json.loads(r.stdout.strip().splitlines()[-1])
    if rooted:
        assert abs(got["heights_value"] - got["value_var"]) <= 1e-9 * abs(got["value_var"])
        gh = np.zeros(S - 1)
        grate = 0.0
        wb = np.asarray(got["blens"])
        for node, par in d["map"][1:]:
            node, par = int(node), int(par)
            gh[par - S - 1] += rate * wb[node - 1]
            if node > S:
                gh[node - S - 1] -= rate * wb[node - 1]
            grate += bl[node - 1] / rate * wb[node - 1]
        np.testing.assert_allclose(got["heights_grad"], gh, rtol=1e-7, atol=1e-6)
        assert got["rate_grad"] == pytest.approx(grate, rel=1e-7)
        # phylo_loglik_heights_autocorr with all substrates equal: identical branch lengths
        assert got["autocorr_value"] == pytest.approx(got["heights_value"], rel=1e-12)
        assert got["autocorr_heights_maxdiff"] <= 1e-6
        assert got["autocorr_rate_grad_sum"] == pytest.approx(grate, rel=1e-7)
        assert got["heights_nops"] == (S - 1) + 1
    want = O.loglik_grad(d["peel"], d["tipmask"], d["weights"], model, bl, subst, fr, rs, ps, rooted=rooted)
    for key in ("value_double", "value_var"):
        assert abs(got[key] - want.logp) <= 1e-10 * abs(want.logp)
    tol = lambda g, w: np.max(np.abs(np.asarray(g) - w) / np.maximum(1, np.abs(w)), initial=0) <= 1e-8
    assert tol(got["blens"], want.grad_blens) and tol(got["blens_mixed"], want.grad_blens)
    assert tol(got["rs"], want.grad_rs) and tol(got["ps"], want.grad_ps)
    nfree = bcount + len(subst) + (4 if model else 0) + 2 * C   # JC69 fixes the frequencies: not an operand
    assert got["n_operands"] == nfree and got["n_operands_mixed"] == bcount   # operands.size() == grads.size()
    if model:
        assert tol(got["subst"], want.grad_subst) and tol(got["freqs"], want.grad_freqs)
    else:
        # the reference's own operator: pruning_loglik(blens) (eigen/prune_stan.hpp:9-17)
        assert abs(got["pruning_loglik"] - want.logp) <= 1e-10 * abs(want.logp)
        assert tol(got["pruning_grad"], want.grad_blens)
        assert got["domain_error"] is True


@pytest.mark.gpu
@pytest.mark.parametrize("devs", ["0", "0,0"])
def test_run_side_publish_creates_the_default_handle(datasets, devs):
    """phylostan_b200.run.publish -- what the patched `phylostan run` calls with its Stan data dict
    (phylostan/phylostan.py:183-287): the handle the shim will find, on one GPU or sharded over several."""
    import argparse
    from oracle import oracle as O
    from phylostan_b200 import encode as E, likelihood as lk, run
    d = datasets["fluA"]
    S, L = d["tipmask"].shape
    tipdata = np.zeros((S, L, 4))
    for s in range(4):
        tipdata[:, :, s] = (d["tipmask"] >> s) & 1
    data = {"peel": d["peel"], "tipdata": tipdata, "weights": d["weights"], "C": 4, "S": S, "L": L}
    arg = argparse.Namespace(model="HKY", clock="strict", gpu_likelihood=True, gpu_devices=devs)
    lik = run.publish(arg, data)
    try:
        assert lk.lib().phylo_b200_get_default() == lik._h.value
        assert lik.info()["shards"] == len(devs.split(","))
        rng = np.random.default_rng(3)
        bl = rng.exponential(0.05, 2 * S - 2) + 1e-4
        kappa, fr, rs, ps = np.array([5.0]), rng.dirichlet(np.ones(4) * 5), E.weibull_rates(0.5, 4), np.full(4, 0.25)
        got = lik.value_grad(bl, kappa, fr, rs, ps)
        want = O.loglik_grad(d["peel"], d["tipmask"], d["weights"], O.HKY, bl, kappa, fr, rs, ps)
        assert abs(got.log_P - want.logp) <= 1e-10 * abs(want.logp)
        assert np.max(np.abs(got.grad_blens - want.grad_blens) / np.maximum(1, np.abs(want.grad_blens))) <= 1e-8
    finally:
        lk.set_default(None)
        lik.close()
