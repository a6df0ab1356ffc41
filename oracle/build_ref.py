"""oracle/_ref: the reference's OWN C++ likelihood, compiled and run here.  TEST INFRASTRUCTURE ONLY.

The reference's eigen/ path bakes tree, alignment, Q and pi into a header rendered from the Jinja2 template
eigen/eigen.j2 (eigen/util.py:104-109) and computes value + branch gradient in `vbsky_loglik` (eigen/eigen.j2:56-168).
This module renders THAT template, read in place from /root/reference (never copied into the repo), with the data of a
test problem, and compiles it with g++ together with oracle/ref_driver.cpp into oracle/_ref/.  Eigen itself is not in
this image; the template is compiled against oracle/mini_eigen, a minimal stand-in for the Eigen operations it uses --
so what runs is the reference's algorithm text, unmodified, on a small matrix class of ours.  The Stan-facing wrapper
eigen/prune_stan.hpp needs Stan Math and is not built; pruner/ does not compile.

Used by tests/golden/make_golden_ref_eigen.py (fixtures for the GPU box, where /root/reference does not exist) and by
__graft_entry__.build() (proves the recipe).  The baked header grows with taxa x sites, so this is for small problems;
bench.py's CPU arm stays the oracle port.
"""
from __future__ import annotations

import os
import subprocess
from typing import List, Sequence, Tuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_EIGEN = "/root/reference/eigen"
OUT = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.exists(os.path.join(REF_EIGEN, "eigen.j2"))


def render(peel: np.ndarray, tipmask: np.ndarray, Q: np.ndarray, pi: np.ndarray) -> str:
    """The template's variables from this repo's encodings.  peel [S-1,3] 1-based post-order (root 2S-1);
    the reference numbers nodes from 0 with the leaves first and the root last (eigen/util.py:60-80), i.e. node k-1."""
    import jinja2
    S, L = tipmask.shape
    nn = 2 * S - 1
    child_parent = [-1] * nn
    for a, b, p in np.asarray(peel, dtype=int):
        child_parent[a - 1] = p - 1
        child_parent[b - 1] = p - 1
    postorder = list(range(S)) + [int(p) - 1 for p in np.asarray(peel, dtype=int)[:, 2]]
    sparse = []
    for s in range(S):
        sparse.append([(x, l, 1.0) for l in range(L) for x in range(4) if (int(tipmask[s, l]) >> x) & 1])
    tpl = jinja2.Template(open(os.path.join(REF_EIGEN, "eigen.j2"), "rt").read())
    return tpl.render(child_parent=child_parent, postorder=postorder, Q=np.asarray(Q, dtype=float),
                      pi=[repr(float(x)) for x in pi], sparse_tip_partials=sparse, num_sites=L)


def build(name: str, peel, tipmask, Q, pi) -> str:
    """Render + compile; returns the path of the executable oracle/_ref/<name>/eigen_ref."""
    d = os.path.join(OUT, name)
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "eigen.hpp"), "wt") as f:
        f.write(render(peel, tipmask, Q, pi))
    exe = os.path.join(d, "eigen_ref")
    cmd = ["g++", "-std=c++14", "-O1", "-ffp-contract=off", "-I", d, "-I", os.path.join(HERE, "mini_eigen"), "-I", REF_EIGEN,
           os.path.join(HERE, "ref_driver.cpp"), "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("compiling the reference's eigen.hpp failed:\n" + r.stderr[-4000:])
    return exe


def run(exe: str, blens_sets: Sequence[Sequence[float]]) -> List[Tuple[float, np.ndarray]]:
    """log_P and the reference's gradient vector (times[i] * dlogP/dtimes[i], eigen.j2:165) for every set of branch lengths."""
    text = "".join(f"{len(t)} " + " ".join(repr(float(x)) for x in t) + "\n" for t in blens_sets)
    r = subprocess.run([exe], input=text, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("the reference binary failed: " + r.stderr[-2000:])
    out = []
    for line in r.stdout.strip().splitlines():
        v = [float(x) for x in line.split()]
        out.append((v[0], np.array(v[1:])))
    return out


def smoke() -> None:
    """The reference's own example (eigen/eigen.cpp:5: every branch length 1) on a three-taxon problem."""
    peel = np.array([[1, 2, 4], [4, 3, 5]])
    tipmask = np.array([[1, 2, 15, 8], [1, 4, 2, 8], [2, 4, 1, 15]], dtype=np.uint8)
    mu = 0.25                                   # eigen/util.py:86-89
    Q = np.full((4, 4), mu)
    np.fill_diagonal(Q, -3 * mu)
    exe = build("smoke", peel, tipmask, Q, np.full(4, 0.25))
    (logp, grad), = run(exe, [[1.0] * 4])
    assert np.isfinite(logp) and logp < 0 and grad.shape == (4,), (logp, grad)


if __name__ == "__main__":
    smoke()
    print("oracle/_ref: reference eigen.j2 rendered, compiled against mini_eigen and run")
