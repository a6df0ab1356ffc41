#!/usr/bin/env python3
"""How much would site-repeat caching save?  (the reference's unfinished pruner/ re-uses a
subtree's partial when the tip states under it equal the previous column's, pruner/tree.cpp:140-174.)

For every data set: (a) the reference's scheme -- share of internal (node, pattern) partials whose subtree tip states
equal those of the PREVIOUS pattern column; (b) the general bound -- share that repeats ANY earlier column's subtree
pattern (distinct subtree patterns per node / L); (c) the same weighted the way this library does the work: per
32 K-pattern warp tile a step can only be skipped when ALL its lanes repeat; (d) cherries (both children tips).
CPU only:   python tools/site_repeats.py > profiles/r2_site_repeats.md
"""
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)


def analyse(name, peel, tipmask, tile=128):
    S, L = tipmask.shape
    cls = {k + 1: tipmask[k].astype(np.int64) for k in range(S)}
    n_int = S - 1
    prev_same = {k + 1: np.concatenate([[False], tipmask[k][1:] == tipmask[k][:-1]]) for k in range(S)}
    tot = rep_prev = rep_any = 0
    tile_skip = tile_tot = 0
    cherries = 0
    rows = []
    for a, b, p in peel:
        a, b, p = int(a), int(b), int(p)
        pair = cls[a] * (int(max(cls[b].max(), 1)) + 1) + cls[b]
        uniq, inv = np.unique(pair, return_inverse=True)
        cls[p] = inv.astype(np.int64)
        prev_same[p] = prev_same[a] & prev_same[b]
        first = np.zeros(L, dtype=bool)
        first[np.unique(inv, return_index=True)[1]] = True
        tot += L
        rep_prev += int(prev_same[p].sum())
        rep_any += int(L - uniq.size)
        # a whole tile of `tile` consecutive patterns is skippable only if none of its patterns is a first occurrence
        nt = (L + tile - 1) // tile
        pad = np.concatenate([first, np.zeros(nt * tile - L, dtype=bool)])
        tile_skip += int((~pad.reshape(nt, tile).any(axis=1)).sum())
        tile_tot += nt
        if a <= S and b <= S:
            cherries += 1
        rows.append(uniq.size)
    return {"name": name, "S": S, "L": L, "prev": rep_prev / tot, "any": rep_any / tot, "tile": tile_skip / tile_tot,
            "cherries": cherries / n_int, "median_distinct": float(np.median(rows)) / L}


def main():
    out = []
    for name in ("fluA", "DS1", "HCV"):
        z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        out.append(analyse(name, z["peel"], z["tipmask"]))
    from phylostan_b200 import synth
    prob = synth.make_problem(1000, 20000, 4, seed=synth.SEED_DATA)
    out.append(analyse("synthetic 1000 taxa x 20k patterns (config-3 generator)", prob.peel, prob.tipmask))
    print("# Site repeats on the reference's data sets (tools/site_repeats.py, CPU)\n")
    print("Share of internal-node partials (node x compressed site pattern) that a site-repeat cache could skip.\n")
    print("| data set | taxa | patterns | equal to the previous column (pruner/tree.cpp:140-174) | repeats of any earlier column "
          "| whole 128-pattern tiles skippable | cherries / internal nodes | median distinct subtree patterns per node / L |")
    print("|---|---|---|---|---|---|---|---|")
    for r in out:
        print(f"| {r['name']} | {r['S']} | {r['L']} | {100 * r['prev']:.1f} % | {100 * r['any']:.1f} % | {100 * r['tile']:.1f} % | "
              f"{100 * r['cherries']:.1f} % | {r['median_distinct']:.3f} |")


def tables():
    """What the library's structural message tables cover, from the tree alone (phylo_b200_plan_tables, CPU)."""
    from phylostan_b200 import likelihood as lk, synth
    rows = []
    for name in ("fluA", "DS1", "HCV"):
        z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        rows.append((name, z["peel"]))
    rows.append(("config-3 tree (1000 taxa)", synth.make_problem(1000, 64, 4, seed=synth.SEED_DATA).peel))
    rows.append(("a 10 000-taxon coalescent tree (config-4 shape)", synth.coalescent_peel_fast(10000, np.random.default_rng(1))))
    print("\n## Structural message tables (DESIGN.md section 7): nodes with at most three tips below them\n")
    print("What the library's message tables cover, from the tree alone (`phylo_b200_plan_tables`, CPU):\n")
    print("| tree | taxa | internal nodes | cherries | pitchforks | share of internal nodes | post-order steps left "
          "| table entries per (draw, category) |")
    print("|---|---|---|---|---|---|---|---|")
    for name, peel in rows:
        S = peel.shape[0] + 1
        t2, t3 = lk.plan_tables(peel, 2), lk.plan_tables(peel, 3)
        c, pf = t2["table_nodes"], t3["table_nodes"] - t2["table_nodes"]
        print(f"| {name} | {S} | {S - 1} | {c} | {pf} | {100 * (c + pf) / (S - 1):.1f} % | {t3['post_steps']} | {t3['table_entries']} |")


if __name__ == "__main__":
    main()
    tables()
