"""Pattern-sharded evaluation across the GPUs of one box: one process per GPU, each holding the full
tree and a contiguous slice of site patterns, one all-reduce of the packed result rows
[logL | d blens | d subst | d freqs | d rs | d ps] per evaluation.

Site patterns are conditionally independent -- ``target += log(...) * weights[i]`` is a plain sum over
``i`` (phylostan/generate_script.py:1010) -- so there is no other exchange on the path.
torch.distributed is the plumbing: NCCL on the device result block in production, gloo on host rows in
the CPU tests.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np


def shard_bounds(L: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced pattern slice [lo, hi) of rank ``rank``; the first L % world ranks get one extra."""
    if not (0 <= rank < world) or L < 0:
        raise ValueError("bad shard request")
    base, extra = divmod(L, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class _CudaView:
    """__cuda_array_interface__ view of the library's device result block."""

    def __init__(self, ptr: int, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def device_out_tensor(lik, B: int):
    """torch tensor aliasing the [B, nout] device result block of a TreeLikelihood (no copy)."""
    import torch
    ptr, ld = lik.device_out()
    return torch.as_tensor(_CudaView(ptr, (B, ld)), device=f"cuda:{lik.device}")


class ShardedLikelihood:
    """Rank-local likelihood + all-reduce.  ``local`` is a TreeLikelihood built on this rank's pattern
    slice (production) or any callable ``(params...) -> packed rows [B, nout]`` (host-side tests)."""

    def __init__(self, local, group=None, layout: Optional[Tuple[int, int, int]] = None):
        """``layout`` = (bcount, nsubst, C) is needed only when ``local`` is a callable."""
        import torch.distributed as dist
        self.local, self.group, self.dist = local, group, dist
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._is_gpu = hasattr(local, "device_out")
        self._stream = None
        if self._is_gpu:
            self.bcount, self.nsubst, self.C = local.bcount, local.nsubst, local.C
            # The library launches on the handle's stream and NCCL on torch's current stream: bind both to ONE
            # stream owned by this object, so the all-reduce can neither overtake the kernels nor be overtaken
            # by the download.  (Whatever stream the caller had set on the handle is replaced.)
            import torch
            self._stream = torch.cuda.Stream(device=local.device)
            local.set_stream(self._stream.cuda_stream)
        elif layout is not None:
            self.bcount, self.nsubst, self.C = layout

    # -- the TreeLikelihood surface the model blocks of phylostan_b200.advi / .sampling call, so that ADVI,
    #    NUTS and lock-step HMC run unchanged over pattern shards (every rank runs the same driver with the
    #    same seed; after the all-reduce every rank holds the same log-likelihood and gradient)
    def _unpack(self, rows: np.ndarray):
        from .likelihood import ValueGrad
        o = 1
        gb = rows[:, o:o + self.bcount]; o += self.bcount
        gs = rows[:, o:o + self.nsubst]; o += self.nsubst
        gf = rows[:, o:o + 4]; o += 4
        gr = rows[:, o:o + self.C]; o += self.C
        return ValueGrad(rows[:, 0].copy(), gb, gs, gf, gr, rows[:, o:o + self.C])

    def value_grad(self, blens, subst=None, freqs=None, rs=None, ps=None):
        blens = np.asarray(blens, dtype=np.float64)
        vg = self._unpack(self.packed(np.atleast_2d(blens), subst, freqs, rs, ps, want_grad=True))
        if blens.ndim == 1:
            return type(vg)(float(vg.log_P[0]), vg.grad_blens[0], vg.grad_subst[0], vg.grad_freqs[0], vg.grad_rs[0],
                            vg.grad_ps[0])
        return vg

    def loglik(self, blens, subst=None, freqs=None, rs=None, ps=None):
        blens = np.asarray(blens, dtype=np.float64)
        lp = self.packed(np.atleast_2d(blens), subst, freqs, rs, ps, want_grad=False)[:, 0].copy()
        return float(lp[0]) if blens.ndim == 1 else lp

    def packed(self, blens, subst=None, freqs=None, rs=None, ps=None, want_grad: bool = True) -> np.ndarray:
        """All-reduced packed rows [B, nout] on every rank."""
        import torch
        if self._is_gpu:
            lik = self.local
            B = lik.upload(blens, subst, freqs, rs, ps)
            lik.run(B, want_grad)
            if self.world > 1:
                with torch.cuda.stream(self._stream):  # the stream the kernels were launched on
                    self.dist.all_reduce(device_out_tensor(lik, B), group=self.group)
            return lik.download(B)                     # copies on, then synchronises, the same stream
        rows = np.atleast_2d(np.asarray(self.local(blens, subst, freqs, rs, ps), dtype=np.float64))
        if not want_grad:
            rows = rows[:, :1]
        if self.world > 1:
            t = torch.from_numpy(rows.copy())
            self.dist.all_reduce(t, group=self.group)
            rows = t.numpy()
        return rows
