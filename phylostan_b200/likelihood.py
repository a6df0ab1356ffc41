"""Python host side of the GPU tree likelihood: a thin ctypes binding of libphylo_b200.so.

Mirrors the reference's operator surface for this path (paths relative to /root/reference):

* ``TreeLikelihood(data...)``       <- the constants eigen/eigen.j2:19-38 bakes into eigen.hpp /
                                       the Stan data dict of phylostan/phylostan.py:183-272
* ``TreeLikelihood.loglik(...)``    <- ``double pruning_loglik(blens, pstream)``  eigen/eigen.j2:171-177
* ``TreeLikelihood.value_grad(...)``<- ``value_grad vbsky_loglik(times)``         eigen/eigen.j2:56-168
                                       (what eigen/prune_stan.hpp:13-16 feeds precomputed_gradients)
* ``pruning_loglik(blens)``         <- the Stan-facing name, eigen/example.stan:3,134

The CUDA extension is the only implementation: if libphylo_b200.so is missing or no sm_100 GPU is
present every entry point raises (there is deliberately no CPU fallback here).
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PHYLO_B200_LIB", os.path.join(_HERE, "csrc", "libphylo_b200.so"))

JC69, HKY, GTR = 0, 1, 2
MODELS = {"JC69": JC69, "HKY": HKY, "GTR": GTR}
N_SUBST = {JC69: 0, HKY: 1, GTR: 6}
ROOTED, NO_NORMQ = 1, 2
EINVAL, ECUDA, ENODEV, EDOMAIN, ENOMEM = -1, -2, -3, -4, -5


class PhyloB200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libphylo_b200 error {code}: {msg}")
        self.code = code


class PhyloDomainError(PhyloB200Error, ValueError):
    """Non-finite / out-of-domain parameters or result (Stan convention: reject the draw)."""


_lib = None
_dp = ctypes.POINTER(ctypes.c_double)


def lib() -> ctypes.CDLL:
    """Load libphylo_b200.so (built in-tree by __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(phylostan_b200 has no CPU fallback)")
    # RTLD_GLOBAL: a pystan-compiled model that includes phylo_b200_stan.hpp binds its undefined
    # phylo_b200_* symbols to this copy at load time (pystan has no link-argument hook)
    L = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    i, vp = ctypes.c_int, ctypes.c_void_p
    ip, bp = ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_uint8)
    L.phylo_b200_create.argtypes = [ctypes.POINTER(vp), i, i, i, i, i, ip, bp, _dp, i]
    L.phylo_b200_create_tipdata.argtypes = [ctypes.POINTER(vp), i, i, i, i, i, ip, _dp, _dp, i]
    if hasattr(L, "phylo_b200_create_device"):  # absent from older A/B builds (tools/ab_sweep.py)
        L.phylo_b200_create_device.argtypes = [ctypes.POINTER(vp), i, i, i, i, i, ip, vp, vp, i]
    if hasattr(L, "phylo_b200_create_multi"):
        L.phylo_b200_create_multi.argtypes = [ctypes.POINTER(vp), i, i, i, i, i, ip, bp, _dp, ip, i]
    L.phylo_b200_destroy.argtypes = [vp]
    L.phylo_b200_destroy.restype = None
    for name in ("bcount", "nsubst", "ncat", "nout", "sync"):
        getattr(L, "phylo_b200_" + name).argtypes = [vp]
    L.phylo_b200_eval.argtypes = [vp, _dp, _dp, _dp, _dp, _dp, i, _dp, _dp, _dp, _dp, _dp, _dp]
    L.phylo_b200_eval_batch.argtypes = [vp, i, _dp, _dp, _dp, _dp, _dp, i, _dp, _dp, _dp, _dp, _dp, _dp]
    L.phylo_b200_eval_heights.argtypes = [vp, ip, _dp, _dp, _dp, i, _dp, _dp, _dp, _dp, i, _dp, _dp, _dp, _dp, _dp,
                                          _dp, _dp]
    L.phylo_b200_eval_heights_autocorr.argtypes = L.phylo_b200_eval_heights.argtypes
    if hasattr(L, "phylo_b200_eval_heights_batch"):
        L.phylo_b200_eval_heights_batch.argtypes = [vp, i, ip, i, _dp, _dp, _dp, i, _dp, _dp, _dp, _dp, i, _dp, _dp, _dp,
                                                    _dp, _dp, _dp, _dp]
        L.phylo_b200_eval_ratios_batch.argtypes = [vp, i, ip, i, _dp, _dp, _dp, _dp, i, _dp, _dp, _dp, _dp, _dp, i, _dp,
                                                   _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp]
    L.phylo_b200_eval_batch_status.argtypes = L.phylo_b200_eval_batch.argtypes + [ip]
    L.phylo_b200_upload.argtypes = [vp, i, _dp, _dp, _dp, _dp, _dp]
    L.phylo_b200_run.argtypes = [vp, i, i]
    L.phylo_b200_device_out.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(i)]
    L.phylo_b200_download.argtypes = [vp, i, _dp]
    L.phylo_b200_set_stream.argtypes = [vp, vp]
    L.phylo_b200_set_tiling.argtypes = [vp, i, i]
    L.phylo_b200_set_precision.argtypes = [vp, i]
    L.phylo_b200_set_stack_slots.argtypes = [vp, i]
    L.phylo_b200_set_sweep_variant.argtypes = [vp, i]
    L.phylo_b200_set_cherry_tables.argtypes = [vp, i]
    L.phylo_b200_set_timing.argtypes = [vp, i]
    L.phylo_b200_get_timing.argtypes = [vp, _dp]
    L.phylo_b200_info.argtypes = [vp, i]
    L.phylo_b200_info.restype = ctypes.c_longlong
    L.phylo_b200_last_error.restype = ctypes.c_char_p
    L.phylo_b200_set_default.argtypes = [vp]
    L.phylo_b200_get_default.restype = vp
    L.phylo_b200_ratios_forward.argtypes = [i, ip, _dp, i, _dp, _dp, _dp, _dp]
    L.phylo_b200_ratios_reverse.argtypes = [i, ip, _dp, i, _dp, _dp, _dp, _dp, _dp]
    L.phylo_b200_plan.argtypes = [i, ip, ip, ip, ip]
    L.phylo_b200_plan_tables.argtypes = [i, ip, i, ip, ip, ip, ip]
    L.phylo_b200_derive.argtypes = [i, i, _dp, _dp, _dp]
    _lib = L
    return L


def _check(rc: int) -> None:
    if rc < 0:
        msg = lib().phylo_b200_last_error().decode()
        raise (PhyloDomainError if rc == EDOMAIN else PhyloB200Error)(rc, msg)


def _arr(a, shape=None) -> Optional[np.ndarray]:
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and a.shape != tuple(shape):
        raise ValueError(f"expected shape {tuple(shape)}, got {a.shape}")
    return a


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(_dp)


@dataclass
class ValueGrad:
    """``struct value_grad`` of eigen/value_grad.hpp:5-8, widened to every differentiable input."""
    log_P: np.ndarray          # [B] (or scalar for a single draw)
    grad_blens: Optional[np.ndarray] = None
    grad_subst: Optional[np.ndarray] = None
    grad_freqs: Optional[np.ndarray] = None
    grad_rs: Optional[np.ndarray] = None
    grad_ps: Optional[np.ndarray] = None

    @property
    def grad(self) -> np.ndarray:
        """Flat gradient in operand order (blens, subst, freqs, rs, ps) -- what the Stan shim hands
        to ``precomputed_gradients`` (eigen/prune_stan.hpp:14-16)."""
        return np.concatenate([np.atleast_1d(g).reshape(-1) for g in
                               (self.grad_blens, self.grad_subst, self.grad_freqs, self.grad_rs, self.grad_ps)
                               if g is not None])


class TreeLikelihood:
    """One tree + alignment resident on one GPU.

    peel [S-1,3] int (1-based, post-order), tipmask [S,L] uint8 (or ``tipdata`` [S,L,4]),
    weights [L]; see phylostan_b200.encode for how these come out of a tree and an alignment.
    """

    def __init__(self, peel, tipmask=None, weights=None, *, tipdata=None, model="GTR", categories: int = 1,
                 rooted: bool = True, normalize: bool = True, device: int = 0, device_tips=None,
                 devices: Optional[Sequence[int]] = None):
        """``device_tips`` = (tipmask_ptr, L, weights_ptr or None): the alignment is already resident on
        ``device`` as uint8 [S, L] masks / float64 [L] weights (e.g. torch tensors' ``data_ptr()``).
        ``devices`` = several CUDA ordinals: ONE handle whose site patterns are sharded over those GPUs
        (``phylo_b200_create_multi``); every call below then runs all shards and returns their sum."""
        L_ = lib()
        self._h = None
        self.model = MODELS[model] if isinstance(model, str) else int(model)
        peel = np.ascontiguousarray(peel, dtype=np.int32)
        self.S = peel.shape[0] + 1
        if peel.shape != (self.S - 1, 3):
            raise ValueError("peel must be [S-1, 3]")
        flags = (ROOTED if rooted else 0) | (0 if normalize else NO_NORMQ)
        self.rooted, self.normalize, self.C, self.device = rooted, normalize, int(categories), int(device)
        ip = ctypes.POINTER(ctypes.c_int32)
        # every shape is checked BEFORE the library reads the arrays
        w = _arr(weights)
        h = ctypes.c_void_p()
        if device_tips is not None:
            tip_ptr, self.L, w_ptr = device_tips
            self.L = int(self.L)
            if self.L < 1 or not tip_ptr:
                raise ValueError("device_tips must be (tipmask_ptr, L >= 1, weights_ptr or None)")
            rc = L_.phylo_b200_create_device(ctypes.byref(h), self.S, self.L, self.C, self.model, flags,
                                             peel.ctypes.data_as(ip), ctypes.c_void_p(int(tip_ptr)),
                                             ctypes.c_void_p(int(w_ptr) if w_ptr else 0), self.device)
        elif tipdata is not None:
            td = _arr(tipdata)
            if td.ndim != 3 or td.shape[0] != self.S or td.shape[2] != 4:
                raise ValueError("tipdata must be [S, L, 4]")
            self.L = td.shape[1]
            if w is not None and w.shape != (self.L,):
                raise ValueError("weights must be [L]")
            rc = L_.phylo_b200_create_tipdata(ctypes.byref(h), self.S, self.L, self.C, self.model, flags,
                                              peel.ctypes.data_as(ip), _ptr(td), _ptr(w), self.device)
        elif devices is not None and len(devices) > 1:
            tm = np.ascontiguousarray(tipmask, dtype=np.uint8)
            if tm.ndim != 2 or tm.shape[0] != self.S:
                raise ValueError("tipmask must be [S, L]")
            self.L = tm.shape[1]
            if w is not None and w.shape != (self.L,):
                raise ValueError("weights must be [L]")
            dv = np.ascontiguousarray(devices, dtype=np.int32)
            self.device = int(dv[0])
            rc = L_.phylo_b200_create_multi(ctypes.byref(h), self.S, self.L, self.C, self.model, flags,
                                            peel.ctypes.data_as(ip),
                                            tm.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), _ptr(w),
                                            dv.ctypes.data_as(ip), len(dv))
        else:
            if devices is not None and len(devices) == 1:
                self.device = int(devices[0])
            tm = np.ascontiguousarray(tipmask, dtype=np.uint8)
            if tm.ndim != 2 or tm.shape[0] != self.S:
                raise ValueError("tipmask must be [S, L]")
            self.L = tm.shape[1]
            if w is not None and w.shape != (self.L,):
                raise ValueError("weights must be [L]")
            rc = L_.phylo_b200_create(ctypes.byref(h), self.S, self.L, self.C, self.model, flags,
                                      peel.ctypes.data_as(ip),
                                      tm.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), _ptr(w), self.device)
        _check(rc)
        self._h = h
        self.bcount = L_.phylo_b200_bcount(h)
        self.nsubst = L_.phylo_b200_nsubst(h)
        self.nout = L_.phylo_b200_nout(h)

    # ------------------------------------------------------------------ lifetime
    def close(self) -> None:
        if getattr(self, "_h", None):
            lib().phylo_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ------------------------------------------------------------------ evaluation
    def _inputs(self, blens, subst, freqs, rs, ps):
        blens = _arr(blens)
        single = blens.ndim == 1
        B = 1 if single else blens.shape[0]
        blens = _arr(blens.reshape(B, -1), (B, self.bcount))
        subst = None if self.nsubst == 0 else _arr(np.reshape(subst, (B, -1)), (B, self.nsubst))
        freqs = None if freqs is None else _arr(np.reshape(freqs, (B, -1)), (B, 4))
        rs = None if rs is None else _arr(np.reshape(rs, (B, -1)), (B, self.C))
        ps = None if ps is None else _arr(np.reshape(ps, (B, -1)), (B, self.C))
        return single, B, blens, subst, freqs, rs, ps

    def value_grad(self, blens, subst=None, freqs=None, rs=None, ps=None) -> ValueGrad:
        """log-likelihood and gradient for one draw ([bcount]) or a batch ([B, bcount])."""
        single, B, blens, subst, freqs, rs, ps = self._inputs(blens, subst, freqs, rs, ps)
        logp = np.zeros(B)
        gb, gs = np.zeros((B, self.bcount)), np.zeros((B, max(self.nsubst, 1)))
        gf, gr, gp = np.zeros((B, 4)), np.zeros((B, self.C)), np.zeros((B, self.C))
        _check(lib().phylo_b200_eval_batch(self._h, B, _ptr(blens), _ptr(subst), _ptr(freqs), _ptr(rs), _ptr(ps), 1,
                                           _ptr(logp), _ptr(gb), _ptr(gs), _ptr(gf), _ptr(gr), _ptr(gp)))
        gs = gs[:, :self.nsubst]
        if single:
            return ValueGrad(float(logp[0]), gb[0], gs[0], gf[0], gr[0], gp[0])
        return ValueGrad(logp, gb, gs, gf, gr, gp)

    def value_grad_masked(self, blens, subst=None, freqs=None, rs=None, ps=None, want_grad=True):
        """A batch in which rejected draws do not fail the call (``phylo_b200_eval_batch_status``): returns
        (ValueGrad, status [B]) with status 0 = evaluated, 1 = parameters out of domain, 2 = log-likelihood not
        finite; rejected draws carry ``-inf`` and a zero gradient -- what Stan does with a rejected draw."""
        _, B, blens, subst, freqs, rs, ps = self._inputs(blens, subst, freqs, rs, ps)
        logp, status = np.zeros(B), np.zeros(B, dtype=np.int32)
        gb, gs = np.zeros((B, self.bcount)), np.zeros((B, max(self.nsubst, 1)))
        gf, gr, gp = np.zeros((B, 4)), np.zeros((B, self.C)), np.zeros((B, self.C))
        g = (_ptr(gb), _ptr(gs), _ptr(gf), _ptr(gr), _ptr(gp)) if want_grad else (None,) * 5
        _check(lib().phylo_b200_eval_batch_status(self._h, B, _ptr(blens), _ptr(subst), _ptr(freqs), _ptr(rs), _ptr(ps),
                                                  int(bool(want_grad)), _ptr(logp), *g,
                                                  status.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))))
        return ValueGrad(logp, gb, gs[:, :self.nsubst], gf, gr, gp), status

    def loglik(self, blens, subst=None, freqs=None, rs=None, ps=None):
        """Value only (post-order sweep only; ADVI's ELBO draws use this)."""
        single, B, blens, subst, freqs, rs, ps = self._inputs(blens, subst, freqs, rs, ps)
        logp = np.zeros(B)
        _check(lib().phylo_b200_eval_batch(self._h, B, _ptr(blens), _ptr(subst), _ptr(freqs), _ptr(rs), _ptr(ps), 0,
                                           _ptr(logp), None, None, None, None, None))
        return float(logp[0]) if single else logp

    def value_grad_heights(self, map_, heights, rates, lowers=None, subst=None, freqs=None, rs=None, ps=None,
                           autocorrelated=False):
        """Clock-tree front end (generate_script.py:660-679): node heights + rate(s) in, gradient with
        respect to heights and rate(s) out.  ``rates``: scalar (strict clock) or [2S-2] substrates.
        ``autocorrelated=True``: branch rate = mean of the rates at both ends (generate_script.py:682-708)."""
        m = np.ascontiguousarray(map_, dtype=np.int32)
        hts = _arr(heights, (self.S - 1,))
        rt = _arr(np.atleast_1d(rates))
        lo = None if lowers is None else _arr(lowers, (2 * self.S - 1,))
        subst = None if self.nsubst == 0 else _arr(subst, (self.nsubst,))
        freqs = None if freqs is None else _arr(freqs, (4,))
        rs = None if rs is None else _arr(rs, (self.C,))
        ps = None if ps is None else _arr(ps, (self.C,))
        logp = np.zeros(1)
        gh, gr = np.zeros(self.S - 1), np.zeros(rt.size)
        gs, gf, grs, gps = np.zeros(max(self.nsubst, 1)), np.zeros(4), np.zeros(self.C), np.zeros(self.C)
        fn = lib().phylo_b200_eval_heights_autocorr if autocorrelated else lib().phylo_b200_eval_heights
        _check(fn(self._h, m.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), _ptr(hts), _ptr(lo), _ptr(rt), rt.size,
                  _ptr(subst), _ptr(freqs), _ptr(rs), _ptr(ps), 1, _ptr(logp), _ptr(gh), _ptr(gr), _ptr(gs), _ptr(gf),
                  _ptr(grs), _ptr(gps)))
        return float(logp[0]), gh, gr, ValueGrad(float(logp[0]), None, gs[:self.nsubst], gf, grs, gps)

    def _clock_inputs(self, B, rates, lowers, subst, freqs, rs, ps):
        rt = _arr(np.reshape(rates, (B, -1)))
        lo = None if lowers is None else _arr(lowers, (2 * self.S - 1,))
        subst = None if self.nsubst == 0 else _arr(np.reshape(subst, (B, -1)), (B, self.nsubst))
        freqs = None if freqs is None else _arr(np.reshape(freqs, (B, -1)), (B, 4))
        rs = None if rs is None else _arr(np.reshape(rs, (B, -1)), (B, self.C))
        ps = None if ps is None else _arr(np.reshape(ps, (B, -1)), (B, self.C))
        return rt, lo, subst, freqs, rs, ps

    def value_grad_heights_batch(self, map_, heights, rates, lowers=None, subst=None, freqs=None, rs=None, ps=None,
                                 autocorrelated=False, want_grad=True):
        """B draws of the clock-tree front end with heights -> blens and its chain rule on the device
        (``phylo_b200_eval_heights_batch``).  heights [B, S-1]; rates [B] / [B, 1] (strict) or [B, 2S-2].
        Returns (logp [B], d/dheights [B, S-1], d/drates [B, nrates], ValueGrad of the site-model parameters)."""
        m = np.ascontiguousarray(map_, dtype=np.int32)
        hts = _arr(np.atleast_2d(heights))
        B = hts.shape[0]
        if hts.shape != (B, self.S - 1):
            raise ValueError("heights must be [B, S-1]")
        rt, lo, subst, freqs, rs, ps = self._clock_inputs(B, rates, lowers, subst, freqs, rs, ps)
        nr = rt.shape[1]
        logp, gh, gr = np.zeros(B), np.zeros((B, self.S - 1)), np.zeros((B, nr))
        gs, gf = np.zeros((B, max(self.nsubst, 1))), np.zeros((B, 4))
        grs, gps = np.zeros((B, self.C)), np.zeros((B, self.C))
        _check(lib().phylo_b200_eval_heights_batch(
            self._h, int(bool(autocorrelated)), m.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), B, _ptr(hts), _ptr(lo),
            _ptr(rt), nr, _ptr(subst), _ptr(freqs), _ptr(rs), _ptr(ps), int(bool(want_grad)), _ptr(logp), _ptr(gh), _ptr(gr),
            _ptr(gs), _ptr(gf), _ptr(grs), _ptr(gps)))
        return logp, gh, gr, ValueGrad(logp, None, gs[:, :self.nsubst], gf, grs, gps)

    def value_grad_ratios_batch(self, map_, props, root_height, rates, lowers=None, subst=None, freqs=None, rs=None,
                                ps=None, hbar_extra=None, autocorrelated=False, want_grad=True, return_heights=True):
        """The same with the ratio parametrisation of the heights (generate_script.py:711-752) on the device too
        (``phylo_b200_eval_ratios_batch``): props [B, S-2], root_height [B]; ``hbar_extra`` [B, S-1] is an extra
        adjoint of the heights (a tree prior's gradient) pushed through the reverse sweep.  Returns a dict:
        logp, logjac, heights, g_props (of logp + logjac), g_root, g_rates, rest (ValueGrad)."""
        m = np.ascontiguousarray(map_, dtype=np.int32)
        pr = _arr(np.atleast_2d(props))
        B = pr.shape[0]
        if pr.shape != (B, self.S - 2):
            raise ValueError("props must be [B, S-2]")
        rh = _arr(np.reshape(root_height, (B,)))
        rt, lo, subst, freqs, rs, ps = self._clock_inputs(B, rates, lowers, subst, freqs, rs, ps)
        hx = None if hbar_extra is None else _arr(np.atleast_2d(hbar_extra), (B, self.S - 1))
        nr = rt.shape[1]
        logp, lj, hts = np.zeros(B), np.zeros(B), (np.zeros((B, self.S - 1)) if return_heights else None)
        gp, groot, gr = np.zeros((B, max(self.S - 2, 1))), np.zeros(B), np.zeros((B, nr))
        gs, gf = np.zeros((B, max(self.nsubst, 1))), np.zeros((B, 4))
        grs, gps = np.zeros((B, self.C)), np.zeros((B, self.C))
        _check(lib().phylo_b200_eval_ratios_batch(
            self._h, int(bool(autocorrelated)), m.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), B, _ptr(pr), _ptr(rh),
            _ptr(lo), _ptr(rt), nr, _ptr(subst), _ptr(freqs), _ptr(rs), _ptr(ps), _ptr(hx), int(bool(want_grad)), _ptr(logp),
            _ptr(lj), _ptr(hts), _ptr(gp), _ptr(groot), _ptr(gr), _ptr(gs), _ptr(gf), _ptr(grs), _ptr(gps)))
        return {"logp": logp, "logjac": lj, "heights": hts, "g_props": gp[:, :self.S - 2], "g_root": groot, "g_rates": gr,
                "rest": ValueGrad(logp, None, gs[:, :self.nsubst], gf, grs, gps)}

    # ------------------------------------------------------------------ resident / split form
    def upload(self, blens, subst=None, freqs=None, rs=None, ps=None) -> int:
        _, B, blens, subst, freqs, rs, ps = self._inputs(blens, subst, freqs, rs, ps)
        _check(lib().phylo_b200_upload(self._h, B, _ptr(blens), _ptr(subst), _ptr(freqs), _ptr(rs), _ptr(ps)))
        return B

    def run(self, B: int, want_grad: bool = True) -> None:
        _check(lib().phylo_b200_run(self._h, int(B), int(bool(want_grad))))

    def download(self, B: int) -> np.ndarray:
        out = np.zeros((B, self.nout))
        _check(lib().phylo_b200_download(self._h, int(B), _ptr(out)))
        return out

    def sync(self) -> None:
        _check(lib().phylo_b200_sync(self._h))

    def device_out(self):
        """(device pointer, leading dimension) of the [B, nout] result buffer."""
        p, ld = ctypes.c_void_p(), ctypes.c_int()
        _check(lib().phylo_b200_device_out(self._h, ctypes.byref(p), ctypes.byref(ld)))
        return p.value, ld.value

    def set_stream(self, stream: Optional[int]) -> None:
        _check(lib().phylo_b200_set_stream(self._h, ctypes.c_void_p(stream or 0)))

    def set_tiling(self, patterns_per_thread: int = 0, pattern_blocks: int = 0) -> None:
        _check(lib().phylo_b200_set_tiling(self._h, patterns_per_thread, pattern_blocks))

    def set_sweep_variant(self, ctas_per_sm: int = -1) -> None:
        """Stack of fp64 K = 4 gradient runs: 0 shared memory, 2 / 3 tensor memory with that many CTAs per SM,
        -1 the library's default (``info()['sweep_variant']`` tells what the last run used)."""
        _check(lib().phylo_b200_set_sweep_variant(self._h, int(ctas_per_sm)))

    def set_cherry_tables(self, enabled: bool = True) -> None:
        """Message-statistic runs take the messages of cherries from 25-entry tables instead of storing and re-reading
        them (``info()['cherry_tables']`` tells whether the last run did)."""
        _check(lib().phylo_b200_set_cherry_tables(self._h, int(bool(enabled))))

    def set_stack_slots(self, slots: int = 0) -> None:
        """Shared-memory stack slots for gradient runs (0 = automatic); fewer than ``info()['stack_depth']``
        parks the top stack positions in the per-CTA HBM scratch."""
        _check(lib().phylo_b200_set_stack_slots(self._h, int(slots)))

    def set_precision(self, bits: int) -> None:
        """64 (default, parity-tested) or 32 (optional fp32-with-scaling mode, error reported separately)."""
        _check(lib().phylo_b200_set_precision(self._h, int(bits)))

    def set_timing(self, enabled: bool) -> None:
        _check(lib().phylo_b200_set_timing(self._h, int(enabled)))

    def get_timing(self) -> dict:
        ms = np.zeros(4)
        _check(lib().phylo_b200_get_timing(self._h, _ptr(ms)))
        return {"pmat_ms": ms[0], "sweep_ms": ms[1], "contract_ms": ms[2], "total_ms": ms[3]}

    def info(self) -> dict:
        names = ["stack_depth", "patterns_per_thread", "threads_per_cta", "grid", "smem_bytes", "padded_patterns",
                 "kernel_launches", "scratch_bytes", "depth_post", "depth_pre", "tiles", "stack_slots", "shards",
                 "sweep_variant", "message_statistic", "cherry_tables", "post_order_tables"]
        return {n: int(lib().phylo_b200_info(self._h, k)) for k, n in enumerate(names)}

    def unpack(self, out: np.ndarray) -> ValueGrad:
        """Split packed rows [B, nout] = [logL | d blens | d subst | d freqs | d rs | d ps]."""
        out = np.atleast_2d(out)
        o = 1
        gb = out[:, o:o + self.bcount]; o += self.bcount
        gs = out[:, o:o + self.nsubst]; o += self.nsubst
        gf = out[:, o:o + 4]; o += 4
        gr = out[:, o:o + self.C]; o += self.C
        gp = out[:, o:o + self.C]
        return ValueGrad(out[:, 0].copy(), gb, gs, gf, gr, gp)


# ---------------------------------------------------------------------- host-only hooks

def plan(peel) -> dict:
    """Traversal plan the library derives from ``peel`` (no GPU needed)."""
    peel = np.ascontiguousarray(peel, dtype=np.int32)
    S = peel.shape[0] + 1
    post = np.zeros((S - 1, 8), dtype=np.int32)
    pre = np.zeros((S - 1, 12), dtype=np.int32)
    depth = np.zeros(2, dtype=np.int32)
    ip = ctypes.POINTER(ctypes.c_int32)
    _check(lib().phylo_b200_plan(S, peel.ctypes.data_as(ip), post.ctypes.data_as(ip), pre.ctypes.data_as(ip),
                                 depth.ctypes.data_as(ip)))
    return {"post": post, "pre": pre, "depth_post": int(depth[0]), "depth_pre": int(depth[1])}


def plan_tables(peel, max_tips: int = 3) -> dict:
    """Message-table nodes of the tree and the second plan whose post-order treats them as leaves (no GPU needed)."""
    peel = np.ascontiguousarray(peel, dtype=np.int32)
    S = peel.shape[0] + 1
    node_tab = np.zeros(2 * S - 1, dtype=np.int32)
    post = np.zeros((S - 1, 8), dtype=np.int32)
    pre = np.zeros((S - 1, 12), dtype=np.int32)
    info = np.zeros(4, dtype=np.int32)
    ip = ctypes.POINTER(ctypes.c_int32)
    _check(lib().phylo_b200_plan_tables(S, peel.ctypes.data_as(ip), int(max_tips), node_tab.ctypes.data_as(ip),
                                        post.ctypes.data_as(ip), pre.ctypes.data_as(ip), info.ctypes.data_as(ip)))
    return {"node_tab": node_tab, "post": post[:info[0]], "pre": pre, "post_steps": int(info[0]), "table_nodes": int(info[1]),
            "table_entries": int(info[2]), "depth": int(info[3])}


def ratios_forward(map_, lowers, props, root_height):
    """heights [B, S-1] and log-Jacobian [B] of the ratio transform (generate_script.py:711-752); no GPU."""
    m = np.ascontiguousarray(map_, dtype=np.int32)
    S = (m.shape[0] + 1) // 2
    props = _arr(np.atleast_2d(props))
    B = props.shape[0]
    root = _arr(np.reshape(root_height, (B,)))
    lo = None if lowers is None else _arr(lowers, (2 * S - 1,))
    heights, logjac = np.empty((B, S - 1)), np.empty(B)
    _check(lib().phylo_b200_ratios_forward(S, m.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), _ptr(lo), B, _ptr(props),
                                           _ptr(root), _ptr(heights), _ptr(logjac)))
    return heights, logjac


def ratios_reverse(map_, lowers, props, heights, hbar):
    """Reverse sweep of ``ratios_forward``: hbar [B, S-1] = d(downstream)/dheights (consumed);
    returns d/dprops [B, S-2] and d/droot_height [B], the log-Jacobian's derivative included."""
    m = np.ascontiguousarray(map_, dtype=np.int32)
    S = (m.shape[0] + 1) // 2
    props, heights = _arr(np.atleast_2d(props)), _arr(np.atleast_2d(heights))
    B = props.shape[0]
    hb = np.array(np.atleast_2d(hbar), dtype=np.float64, order="C")
    lo = None if lowers is None else _arr(lowers, (2 * S - 1,))
    gp, gr = np.empty((B, S - 2)), np.empty(B)
    _check(lib().phylo_b200_ratios_reverse(S, m.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), _ptr(lo), B, _ptr(props),
                                           _ptr(heights), _ptr(hb), _ptr(gp), _ptr(gr)))
    return gp, gr


def derive(model, subst=None, freqs=None, normalize: bool = True) -> dict:
    """Per-draw model algebra (Q, eigen-system, X_theta) computed by the library (no GPU needed)."""
    model = MODELS[model] if isinstance(model, str) else int(model)
    out = np.zeros(56 + 160)
    s = _arr(np.zeros(1) if subst is None else np.atleast_1d(subst))
    f = _arr(np.full(4, 0.25) if freqs is None else freqs)
    nt = lib().phylo_b200_derive(model, 0 if normalize else NO_NORMQ, _ptr(s), _ptr(f), _ptr(out))
    _check(nt)
    return {"pi": out[0:4], "lam": out[4:8], "m1": out[8:24].reshape(4, 4), "m2": out[24:40].reshape(4, 4),
            "Q": out[40:56].reshape(4, 4), "X": out[56:56 + 16 * nt].reshape(nt, 4, 4)}


# ---------------------------------------------------------------------- eigen/ operator surface

_default: Optional[TreeLikelihood] = None


def set_default(lik: TreeLikelihood) -> None:
    """Install the tree/alignment that ``pruning_loglik`` evaluates -- the run-time analogue of
    rendering eigen/eigen.j2 into eigen.hpp (eigen/util.py:104-109)."""
    global _default
    _default = lik
    lib().phylo_b200_set_default(lik._h if lik is not None else None)


def pruning_loglik(blens: Sequence[float]) -> float:
    """``real pruning_loglik(vector blens)`` (eigen/example.stan:3, eigen/eigen.j2:171-177)."""
    if _default is None:
        raise RuntimeError("call set_default(TreeLikelihood(...)) first")
    return _default.loglik(blens)


def pruning_loglik_value_grad(blens: Sequence[float]) -> ValueGrad:
    """What eigen/prune_stan.hpp:13-16 computes before calling precomputed_gradients."""
    if _default is None:
        raise RuntimeError("call set_default(TreeLikelihood(...)) first")
    return _default.value_grad(blens)
