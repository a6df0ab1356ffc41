"""Synthetic trees, alignments and parameter draws of the shapes BASELINE.json names
(SURVEY.md section 8d): random Kingman-coalescent topology, GTR + Weibull(4) simulated columns,
1 % ambiguous tip cells, weights 1 + Poisson(1).  Bench / test harness support, numpy only.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple

import numpy as np

from .encode import weibull_rates

SEED_DATA = 20261018
SEED_DRAWS = 20261019
RATES0 = np.array([0.10, 0.30, 0.10, 0.10, 0.30, 0.10])
FREQS0 = np.array([0.30, 0.20, 0.20, 0.30])
WSHAPE0 = 0.5


def coalescent_peel(S: int, rng: np.random.Generator) -> np.ndarray:
    """Random rooted binary tree (uniform pair merges), returned as the reference's ``peel``:
    tips 1..S, internals S+1..2S-1 numbered in depth-first post-order, root 2S-1
    (phylostan/utils.py:59-81)."""
    left = {}
    right = {}
    active = list(range(S))
    nxt = S
    for _ in range(S - 1):
        i, j = rng.choice(len(active), size=2, replace=False)
        a, b = active[i], active[j]
        left[nxt], right[nxt] = a, b
        active = [x for k, x in enumerate(active) if k != i and k != j] + [nxt]
        nxt += 1
    root = nxt - 1
    # renumber internals in DFS post-order (iterative)
    new_id = {}
    rows = []
    counter = S
    stack = [(root, 0)]
    while stack:
        node, stage = stack.pop()
        if node < S:
            new_id[node] = node
            continue
        if stage == 0:
            stack.append((node, 1))
            stack.append((right[node], 0))
            stack.append((left[node], 0))
        else:
            new_id[node] = counter
            counter += 1
            rows.append((new_id[left[node]] + 1, new_id[right[node]] + 1, new_id[node] + 1))
    return np.asarray(rows, dtype=np.int32)


def coalescent_peel_fast(S: int, rng: np.random.Generator) -> np.ndarray:
    """Same distribution as ``coalescent_peel`` in O(S) list work (swap-remove instead of rebuilding the active
    list), for the 10 000-taxon shape.  Not the same tree for the same seed."""
    left, right = {}, {}
    active = list(range(S))
    nxt = S
    for _ in range(S - 1):
        n = len(active)
        i = int(rng.integers(n))
        j = int(rng.integers(n - 1))
        if j >= i:
            j += 1
        a, b = active[i], active[j]
        left[nxt], right[nxt] = a, b
        for k in sorted((i, j), reverse=True):
            active[k] = active[-1]
            active.pop()
        active.append(nxt)
        nxt += 1
    root = nxt - 1
    new_id, rows, counter = {}, [], S
    stack = [(root, 0)]
    while stack:
        node, stage = stack.pop()
        if node < S:
            new_id[node] = node
            continue
        if stage == 0:
            stack.append((node, 1))
            stack.append((right[node], 0))
            stack.append((left[node], 0))
        else:
            new_id[node] = counter
            counter += 1
            rows.append((new_id[left[node]] + 1, new_id[right[node]] + 1, new_id[node] + 1))
    return np.asarray(rows, dtype=np.int32)


def gtr_q(rates: np.ndarray, freqs: np.ndarray) -> np.ndarray:
    R = np.zeros((4, 4))
    R[np.triu_indices(4, 1)] = rates
    R = R + R.T
    Q = R * freqs[None, :]
    np.fill_diagonal(Q, -Q.sum(1))
    return Q / -(np.diag(Q) * freqs).sum()


def _pmat(Q: np.ndarray, freqs: np.ndarray, tau: float) -> np.ndarray:
    sq = np.sqrt(freqs)
    A = sq[:, None] * Q / sq[None, :]
    lam, U = np.linalg.eigh((A + A.T) / 2)
    return ((U / sq[:, None]) * np.exp(lam * tau)[None, :]) @ (U.T * sq[None, :])


def simulate_alignment(peel: np.ndarray, blens: np.ndarray, L: int, C: int, rng: np.random.Generator,
                       rates=RATES0, freqs=FREQS0, wshape=WSHAPE0, ambiguous: float = 0.01,
                       structured: bool = True) -> Tuple[np.ndarray, np.ndarray]:
    """One independent column per pattern.  Returns (tipmask uint8 [S, L], weights [L])."""
    S = peel.shape[0] + 1
    if structured:
        freqs = np.asarray(freqs)
        Q = gtr_q(np.asarray(rates), freqs)
        sq = np.sqrt(freqs)
        A = sq[:, None] * Q / sq[None, :]
        lam, U = np.linalg.eigh((A + A.T) / 2)
        m1, m2 = U / sq[:, None], U.T * sq[None, :]
        rs = weibull_rates(wshape, C)
        cat = rng.integers(0, C, size=L)
        state = np.zeros((2 * S - 1, L), dtype=np.int8)
        state[2 * S - 2] = rng.choice(4, size=L, p=freqs)
        for row in peel[::-1]:  # parents before children
            par = state[row[2] - 1]
            for child in (row[0] - 1, row[1] - 1):
                P = np.einsum("ik,ck,kj->cij", m1, np.exp(lam[None, :] * (blens[child] * rs)[:, None]), m2)
                cum = np.cumsum(P, axis=2)[:, :, :3]                     # [C, 4, 3]
                u = rng.random(L)
                state[child] = (u[:, None] > cum[cat, par]).sum(1)
        tips = state[:S]
    else:  # cheap generator for very large shapes: iid states biased towards a per-column majority
        major = rng.integers(0, 4, size=L, dtype=np.int8)
        tips = np.where(rng.random((S, L)) < 0.8, major[None, :], rng.integers(0, 4, size=(S, L), dtype=np.int8))
    tipmask = (1 << tips.astype(np.uint8)).astype(np.uint8)
    tipmask[rng.random((S, L)) < ambiguous] = 0xF
    weights = 1.0 + rng.poisson(1.0, size=L)
    return tipmask, weights.astype(np.float64)


def simulate_alignment_device(peel: np.ndarray, blens: np.ndarray, L: int, C: int, device, seed: int,
                              rates=RATES0, freqs=FREQS0, wshape=WSHAPE0, ambiguous: float = 0.01):
    """``simulate_alignment`` on a CUDA device with torch (same model: one independent GTR + Weibull(C) column
    per pattern, ``ambiguous`` of the tip cells set to all ones, weights 1 + Poisson(1)); torch's generator, so
    not the same columns as the numpy version.  Returns device tensors (tipmask uint8 [S, L], weights float64
    [L]) -- the alignment never touches the host, which is what the 10 000 x 1 000 000 shape needs."""
    import torch
    S = peel.shape[0] + 1
    freqs = np.asarray(freqs)
    Q = gtr_q(np.asarray(rates), freqs)
    sq = np.sqrt(freqs)
    A = sq[:, None] * Q / sq[None, :]
    lam, U = np.linalg.eigh((A + A.T) / 2)
    m1, m2 = U / sq[:, None], U.T * sq[None, :]
    rs = weibull_rates(wshape, C)
    tau = blens[:, None] * rs[None, :]                                              # [2S-2, C]
    P = np.einsum("ik,bck,kj->bcij", m1, np.exp(lam[None, None, :] * tau[:, :, None]), m2)
    cum = np.cumsum(P, axis=3)[:, :, :, :3].reshape(2 * S - 2, C * 4, 3)              # [branch, cat*4 + parent, 3]
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    cum_d = torch.from_numpy(np.ascontiguousarray(cum)).to(device)
    cat4 = torch.randint(0, C, (L,), generator=g, device=device) * 4
    tips = torch.empty((S, L), dtype=torch.uint8, device=device)
    internal = torch.empty((S - 1, L), dtype=torch.uint8, device=device)               # states of nodes S .. 2S-2
    pi_cum = torch.from_numpy(np.cumsum(freqs)[:3].copy()).to(device)
    internal[S - 2] = (torch.rand(L, generator=g, device=device, dtype=torch.float64)[:, None] > pi_cum).sum(1).to(torch.uint8)
    for row in peel[::-1]:  # parents before children
        par = internal[row[2] - 1 - S].long() + cat4
        for child in (int(row[0]) - 1, int(row[1]) - 1):
            u = torch.rand(L, generator=g, device=device, dtype=torch.float64)
            st = (u[:, None] > cum_d[child][par]).sum(1).to(torch.uint8)
            if child < S:
                tips[child] = st
            else:
                internal[child - S] = st
    del internal
    tipmask = torch.bitwise_left_shift(torch.ones((), dtype=torch.uint8, device=device), tips)
    del tips
    for s0 in range(0, S, 256):   # ambiguity in slabs: no [S, L] float temporary
        blk = tipmask[s0:s0 + 256]
        blk[torch.rand(blk.shape, generator=g, device=device) < ambiguous] = 0xF
    weights = 1.0 + torch.poisson(torch.ones(L, device=device, dtype=torch.float64), generator=g)
    return tipmask, weights


@dataclass
class SynthProblem:
    S: int
    L: int
    C: int
    peel: np.ndarray
    tipmask: np.ndarray
    weights: np.ndarray
    blens: np.ndarray   # [2S-2] base branch lengths (rooted convention)


def make_problem(S: int, L: int, C: int = 4, seed: int = SEED_DATA, structured: bool = True) -> SynthProblem:
    rng = np.random.default_rng(seed)
    peel = coalescent_peel(S, rng)
    blens = np.clip(rng.exponential(0.02, size=2 * S - 2), 1e-4, 0.5)
    tipmask, weights = simulate_alignment(peel, blens, L, C, rng, structured=structured)
    return SynthProblem(S, L, C, peel, tipmask, weights, blens)


def make_draws(prob: SynthProblem, B: int, seed: int = SEED_DRAWS):
    """B parameter draws around the simulation truth: blens * LogNormal(0, 0.1), Dirichlet(200 x)
    rates and frequencies, wshape * LogNormal(0, 0.1).  Returns (blens, rates, freqs, rs, ps)."""
    rng = np.random.default_rng(seed)
    bl = prob.blens[None, :] * rng.lognormal(0.0, 0.1, size=(B, prob.blens.size))
    rates = rng.dirichlet(200.0 * RATES0 / RATES0.sum(), size=B)
    freqs = rng.dirichlet(200.0 * FREQS0, size=B)
    wshape = WSHAPE0 * rng.lognormal(0.0, 0.1, size=B)
    rs = np.stack([weibull_rates(w, prob.C) for w in wshape])
    ps = np.full((B, prob.C), 1.0 / prob.C)
    return bl, rates, freqs, rs, ps
