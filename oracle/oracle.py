"""ctypes front-end of the CPU ORACLE (oracle/phylo_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under phylostan_b200/ imports this.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libphylo_oracle.so")

JC69, HKY, GTR = 0, 1, 2
ROOTED, NO_NORMQ, RESCALE, QUIRK_TIMES, DP_EIGEN = 1, 2, 4, 8, 16
MODEL_IDS = {"JC69": JC69, "HKY": HKY, "GTR": GTR}
N_SUBST = {JC69: 0, HKY: 1, GTR: 6}


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc, OpenMP when the toolchain has it)."""
    src = os.path.join(_HERE, "phylo_oracle.c")
    if not force and os.path.exists(_LIB_PATH) and os.path.getmtime(_LIB_PATH) >= max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "phylo_oracle.h"))):
        return _LIB_PATH
    base = ["-O3", "-march=x86-64-v3", "-fPIC", "-std=c11", "-fno-fast-math", "-ffp-contract=off",
            "-shared", "-o", _LIB_PATH, src, "-lm"]
    errs = []
    for cc in ("/usr/bin/gcc", "gcc", "cc"):
        for omp in (["-fopenmp"], []):
            r = subprocess.run([cc] + omp + base, capture_output=True, text=True)
            if r.returncode == 0:
                return _LIB_PATH
            errs.append(r.stderr[-300:])
    raise RuntimeError("could not build the oracle: " + " | ".join(errs))


_lib = None


def lib():
    global _lib
    if _lib is None:
        try:
            build()
            _lib = ctypes.CDLL(_LIB_PATH)
        except OSError:
            build(force=True)
            _lib = ctypes.CDLL(_LIB_PATH)
        dp, ip, bp = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_uint8)
        i = ctypes.c_int
        _lib.oracle_loglik_grad.argtypes = [i, i, i, ip, bp, dp, i, i, dp, dp, dp, dp, dp, i, i, dp, dp, dp, dp, dp, dp]
        _lib.oracle_loglik_grad.restype = i
        _lib.oracle_site_loglik.argtypes = [i, i, i, ip, bp, i, i, dp, dp, dp, dp, dp, dp]
        _lib.oracle_site_loglik.restype = i
        _lib.oracle_pmatrix.argtypes = [i, i, dp, dp, ctypes.c_double, dp]
        _lib.oracle_pmatrix.restype = i
        _lib.oracle_pq_invariant.argtypes = [i, i, i, ip, bp, i, i, dp, dp, dp, dp, dp, i, i]
        _lib.oracle_pq_invariant.restype = ctypes.c_double
        _lib.oracle_num_threads.restype = i
    return _lib


def _d(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _p(a, t=ctypes.c_double):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(t))


@dataclass
class Result:
    logp: float
    grad_blens: Optional[np.ndarray] = None
    grad_subst: Optional[np.ndarray] = None
    grad_freqs: Optional[np.ndarray] = None
    grad_rs: Optional[np.ndarray] = None
    grad_ps: Optional[np.ndarray] = None

    def flat(self) -> np.ndarray:
        parts = [np.atleast_1d(self.logp)]
        for g in (self.grad_blens, self.grad_subst, self.grad_freqs, self.grad_rs, self.grad_ps):
            if g is not None:
                parts.append(g)
        return np.concatenate(parts)


def _flags(rooted, normalize, rescale, quirk_times=False, dp_eigen=False):
    return ((ROOTED if rooted else 0) | (0 if normalize else NO_NORMQ) | (RESCALE if rescale else 0)
            | (QUIRK_TIMES if quirk_times else 0) | (DP_EIGEN if dp_eigen else 0))


def _prep(peel, tipmask, weights, blens, subst, freqs, rs, ps, model):
    peel = np.ascontiguousarray(peel, dtype=np.int32)
    tipmask = np.ascontiguousarray(tipmask, dtype=np.uint8)
    S, L = tipmask.shape
    assert peel.shape == (S - 1, 3)
    rs = _d(np.ones(1) if rs is None else rs)
    ps = _d(np.ones(1) if ps is None else ps)
    freqs = _d(np.full(4, 0.25) if freqs is None else freqs)
    subst = _d(np.zeros(1) if subst is None or np.size(subst) == 0 else np.atleast_1d(subst))
    assert subst.size >= max(1, N_SUBST[model])
    return peel, tipmask, _d(weights), _d(blens), subst, freqs, rs, ps, S, L, rs.size


def loglik_grad(peel, tipmask, weights, model, blens, subst=None, freqs=None, rs=None, ps=None, *,
                rooted=True, normalize=True, rescale=True, quirk_times=False, dp_eigen=False,
                want_grad=True, nthreads=0) -> Result:
    peel, tipmask, weights, blens, subst, freqs, rs, ps, S, L, C = _prep(
        peel, tipmask, weights, blens, subst, freqs, rs, ps, model)
    bcount = 2 * S - 2 if rooted else 2 * S - 3
    assert blens.size == bcount, (blens.size, bcount)
    logp = np.zeros(1)
    gb, gs, gf = np.zeros(bcount), np.zeros(max(1, N_SUBST[model])), np.zeros(4)
    gr, gp = np.zeros(C), np.zeros(C)
    rc = lib().oracle_loglik_grad(S, L, C, _p(peel, ctypes.c_int32), _p(tipmask, ctypes.c_uint8), _p(weights),
                                  model, _flags(rooted, normalize, rescale, quirk_times, dp_eigen), _p(blens),
                                  _p(subst), _p(freqs), _p(rs), _p(ps), int(want_grad), nthreads, _p(logp),
                                  _p(gb), _p(gs), _p(gf), _p(gr), _p(gp))
    if rc:
        raise ValueError(f"oracle_loglik_grad failed: {rc}")
    if not want_grad:
        return Result(float(logp[0]))
    return Result(float(logp[0]), gb, gs[:N_SUBST[model]], gf, gr, gp)


def site_loglik(peel, tipmask, model, blens, subst=None, freqs=None, rs=None, ps=None, *, rooted=True,
                normalize=True, rescale=True) -> np.ndarray:
    peel, tipmask, _, blens, subst, freqs, rs, ps, S, L, C = _prep(
        peel, tipmask, None, blens, subst, freqs, rs, ps, model)
    out = np.zeros(L)
    rc = lib().oracle_site_loglik(S, L, C, _p(peel, ctypes.c_int32), _p(tipmask, ctypes.c_uint8), model,
                                  _flags(rooted, normalize, rescale), _p(blens), _p(subst), _p(freqs), _p(rs),
                                  _p(ps), _p(out))
    if rc:
        raise ValueError(f"oracle_site_loglik failed: {rc}")
    return out


def pmatrix(model, tau, subst=None, freqs=None, normalize=True) -> np.ndarray:
    subst = _d(np.zeros(1) if subst is None else np.atleast_1d(subst))
    freqs = _d(np.full(4, 0.25) if freqs is None else freqs)
    P = np.zeros(16)
    lib().oracle_pmatrix(model, _flags(True, normalize, False), _p(subst), _p(freqs), float(tau), _p(P))
    return P.reshape(4, 4)


def pq_invariant(peel, tipmask, model, blens, subst=None, freqs=None, rs=None, ps=None, *, rooted=True,
                 normalize=True, l=0, c=0) -> float:
    peel, tipmask, _, blens, subst, freqs, rs, ps, S, L, C = _prep(
        peel, tipmask, None, blens, subst, freqs, rs, ps, model)
    return float(lib().oracle_pq_invariant(S, L, C, _p(peel, ctypes.c_int32), _p(tipmask, ctypes.c_uint8), model,
                                           _flags(rooted, normalize, False), _p(blens), _p(subst), _p(freqs),
                                           _p(rs), _p(ps), l, c))


def num_threads() -> int:
    return int(lib().oracle_num_threads())
