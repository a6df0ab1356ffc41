"""integration/phylostan-gpu-likelihood.patch is an artefact a phylostan maintainer can apply: it must apply to the
reference's phylostan/phylostan.py (:11-12, :69-95, :149-159, :292-300) and the patched command line must build the
external-likelihood program.  Needs /root/reference (absent on the GPU box: skipped there)."""
import os
import shutil
import subprocess
import sys
import textwrap

import pytest

from conftest import ROOT

REF = "/root/reference"
PATCH = os.path.join(ROOT, "integration", "phylostan-gpu-likelihood.patch")
needs_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "phylostan")), reason="reference checkout not mounted")


def test_patch_file_is_shipped_and_touches_only_the_command_line():
    txt = open(PATCH).read()
    assert txt.startswith("--- a/phylostan/phylostan.py\n+++ b/phylostan/phylostan.py\n")
    assert txt.count("\n--- ") == 0 and "gpu_run.publish(arg, data)" in txt and "gpu_run.build_script(arg)" in txt


@needs_ref
def test_patch_applies_to_the_reference(tmp_path):
    shutil.copytree(os.path.join(REF, "phylostan"), tmp_path / "phylostan")
    dry = subprocess.run(["patch", "--dry-run", "-p1", "-d", str(tmp_path), "-i", PATCH], capture_output=True, text=True)
    assert dry.returncode == 0 and "FAILED" not in dry.stdout and "fuzz" not in dry.stdout, dry.stdout + dry.stderr
    real = subprocess.run(["patch", "-p1", "-d", str(tmp_path), "-i", PATCH], capture_output=True, text=True)
    assert real.returncode == 0, real.stdout + real.stderr
    subprocess.run([sys.executable, "-m", "py_compile", str(tmp_path / "phylostan" / "phylostan.py")], check=True)


@needs_ref
@pytest.mark.parametrize("flags,expect", [
    (["-m", "HKY", "-C", "4", "--gpu-likelihood"], "target += phylo_loglik(blens, rep_vector(kappa, 1), freqs, rs, ps);"),
    (["-m", "GTR", "-C", "4", "--clock", "strict", "--heterochronous", "--estimate_rate", "-c", "constant", "--gpu-likelihood",
      "--gpu-heights"], "target += phylo_loglik_heights(heights, rep_array(rate, 1), map, lowers, rates, freqs, rs, ps);"),
    (["-m", "GTR", "-C", "4"], "calculate_gtr_p_matrices"),          # without the flag: the stock program
])
def test_patched_command_line_builds_the_external_program(tmp_path, flags, expect):
    """`phylostan build` of the patched checkout, through phylostan_b200.run.patched_checkout; pystan and dendropy
    (absent here, imported at the top of phylostan.py) are stubbed -- `build` uses neither."""
    stubs = tmp_path / "stubs"
    stubs.mkdir()
    (stubs / "pystan.py").write_text("class StanModel:\n    def __init__(self, **kw): raise RuntimeError('no pystan here')\n")
    (stubs / "dendropy.py").write_text("class Tree: pass\nclass DnaCharacterMatrix: pass\nclass TaxonNamespace: pass\n")
    script = tmp_path / "model.stan"
    code = textwrap.dedent(f"""
        import sys
        sys.path[:0] = [{str(stubs)!r}, {ROOT!r}]
        from phylostan_b200 import run
        sys.path.insert(0, run.patched_checkout({str(tmp_path / 'co')!r}, {REF!r}))
        import phylostan.phylostan as ps
        assert ps.gpu_run is run
        sys.argv = ['phylostan', 'build', '-s', {str(script)!r}] + {flags!r}
        ps.main()
    """)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    txt = script.read_text()
    assert expect in txt
    if "--gpu-likelihood" in flags:
        assert "calculate_" not in txt and "partials[" not in txt


def test_run_side_helpers_without_gpu():
    import argparse
    from phylostan_b200 import run
    p = argparse.ArgumentParser()
    run.add_arguments(p)
    a = p.parse_args(["--gpu-likelihood", "--gpu-devices", "0,1,3"])
    assert run.enabled(a) and run.devices(a) == [0, 1, 3]
    assert not run.enabled(p.parse_args([])) and run.devices(p.parse_args([])) == [0]
    with pytest.raises(ValueError):
        run.devices(p.parse_args(["--gpu-devices", "-1"]))


def test_stan_model_hook_loads_the_library_first_and_passes_the_shim(tmp_path):
    """run.stan_model = what the patched phylostan calls instead of pystan.StanModel(file=...): libphylo_b200 is loaded
    RTLD_GLOBAL before the compile (so the model extension's phylo_b200_* symbols bind), and the compile gets
    allow_undefined + the shim header + include directories that exist (the eigen/eigen.py:79-87 mechanism)."""
    from phylostan_b200 import likelihood, run
    seen = {}

    class FakePystan:
        @staticmethod
        def StanModel(file=None, **kw):
            seen["lib_loaded"] = likelihood._lib is not None
            seen["file"], seen["kw"] = file, kw
            return "model"
    script = tmp_path / "m.stan"
    script.write_text("model{}")
    assert run.stan_model(str(script), FakePystan) == "model"
    assert seen["lib_loaded"] and seen["file"] == str(script)
    kw = seen["kw"]
    assert kw["allow_undefined"] is True and kw["includes"] == ["phylo_b200_stan.hpp"]
    assert all(os.path.isdir(d) for d in kw["include_dirs"])
    assert any(os.path.exists(os.path.join(d, "phylo_b200_stan.hpp")) for d in kw["include_dirs"])
    assert any(os.path.exists(os.path.join(d, "phylo_b200.h")) for d in kw["include_dirs"])
