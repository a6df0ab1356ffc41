#!/usr/bin/env python3
"""A/B timing of sweep-kernel builds on the BASELINE config-3 shape (GPU box only).

    python tools/ab_sweep.py [B] [L] [S]            # parent: one child process per library
    PHYLO_AB_LIBS="r1,nopf,..."                      # names under phylostan_b200/csrc/variants/, "" = the in-tree build;
                                                     # "name@3" runs it with PHYLO_B200_SWEEP_TM=3 (tensor-memory stack)
    PHYLO_AB_CASES="4:0,4:5,2:0"                     # K:slots pairs (slots 0 = automatic)

Every child prints the sweep time of each case (CUDA events of the library, best of 3 after 2 warm-ups) and
the results of draw 0; the parent compares draw 0 across libraries (the r1 build is the parity-tested
reference of this comparison) at the north-star tolerances.
"""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)


def child(B, L, S, cases, grad_modes):
    from phylostan_b200 import likelihood as lk, synth
    prob = synth.make_problem(S, L, 4, structured=False)
    draws = synth.make_draws(prob, B)
    lik = lk.TreeLikelihood(prob.peel, prob.tipmask, prob.weights, model="GTR", categories=4)
    lik.upload(*draws)
    res = []
    for grad in grad_modes:
        for K, slots in cases:
            lik.set_tiling(K, 1)
            lik.set_stack_slots(slots)
            try:
                lik.set_timing(False)
                for _ in range(2):
                    lik.run(B, grad)
                lik.sync()
                lik.set_timing(True)
                ms = []
                for _ in range(3):
                    lik.run(B, grad)
                    ms.append(lik.get_timing()["sweep_ms"])
                out = lik.download(B)
                info = lik.info()
                res.append({"grad": grad, "K": K, "slots_req": slots, "slots": info["stack_slots"], "depth": info["stack_depth"],
                            "grid": info["grid"], "smem": info["smem_bytes"], "sweep_ms": min(ms),
                            "evals_per_s": B / min(ms) * 1e3, "row0": out[0].tolist()})
            except Exception as e:  # a variant that does not fit is a result too
                res.append({"grad": grad, "K": K, "slots_req": slots, "error": str(e)})
    print("ABRESULT " + json.dumps(res), flush=True)


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
    S = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
    cases = [tuple(int(x) for x in c.split(":")) for c in os.environ.get("PHYLO_AB_CASES", "4:0,2:0").split(",")]
    grad_modes = [True] if os.environ.get("PHYLO_AB_GRAD_ONLY") else [True, False]
    if os.environ.get("PHYLO_AB_CHILD"):
        return child(B, L, S, cases, grad_modes)
    libs = os.environ.get("PHYLO_AB_LIBS", "r1,").split(",")
    ref = None
    for name in libs:
        env = dict(os.environ, PHYLO_AB_CHILD="1")
        name, _, tm = name.partition("@")   # "lib@3": run that library with PHYLO_B200_SWEEP_TM=3
        if tm.startswith("m"):               # "lib@m0": the same with PHYLO_B200_MSG=0 (no message statistic)
            env["PHYLO_B200_MSG"] = tm[1:]
        elif tm.startswith("c"):             # "lib@c0": PHYLO_B200_CHERRY=0 (no cherry tables)
            env["PHYLO_B200_CHERRY"] = tm[1:]
        elif tm:
            env["PHYLO_B200_SWEEP_TM"] = tm
        if name:
            env["PHYLO_B200_LIB"] = os.path.join(ROOT, "phylostan_b200", "csrc", "variants", f"libphylo_b200_{name}.so")
        p = subprocess.run([sys.executable, os.path.abspath(__file__)] + sys.argv[1:], env=env, capture_output=True, text=True)
        line = [l for l in p.stdout.splitlines() if l.startswith("ABRESULT ")]
        if not line:
            print(f"[{name or 'in-tree'}] FAILED rc={p.returncode}\n{p.stdout[-2000:]}\n{p.stderr[-3000:]}", flush=True)
            continue
        for r in json.loads(line[0][9:]):
            tag = f"[{(name or 'in-tree') + ('@' + tm if tm else ''):10s}] grad={int(r['grad'])} K={r['K']} slots={r.get('slots', '-')}/{r.get('depth', '-')}"
            if "error" in r:
                print(f"{tag} ERROR {r['error']}", flush=True)
                continue
            row = np.array(r.pop("row0"))
            key = (r["grad"],)
            if ref is None:
                ref = {}
            if key not in ref:
                ref[key] = row
            want = ref[key]
            n = row.size if r["grad"] else 1
            e_l = abs(row[0] - want[0]) / abs(want[0])
            e_g = float(np.max(np.abs(row[1:n] - want[1:n]) / np.maximum(1.0, np.abs(want[1:n])))) if n > 1 else 0.0
            ok = e_l <= 1e-10 and e_g <= 1e-8
            print(f"{tag} grid={r['grid']} smem={r['smem']} sweep={r['sweep_ms']:.2f} ms -> {r['evals_per_s']:.1f} evals/s  "
                  f"vs first: logL rel {e_l:.1e} grad {e_g:.1e} {'ok' if ok else 'MISMATCH'}", flush=True)


if __name__ == "__main__":
    main()
