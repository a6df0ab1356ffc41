import sys, os; sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import numpy as np
from oracle import oracle as O
from phylostan_b200 import likelihood as lk
import test_gpu_parity as T
for name, model, C in [("DS1", O.GTR, 4), ("fluA", O.HKY, 4), ("HCV", O.GTR, 1), ("DS1", O.JC69, 3)]:
    d = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", name + ".npz"))
    rooted = bool(d["rooted"]); S = d["tipmask"].shape[0]
    rng = np.random.default_rng(101)
    draws = T._extreme_draws(model, S, rooted, C, rng, 12)
    with T.make(d["peel"], d["tipmask"], d["weights"], model, C, rooted=rooted) as lik:
        for i, (bl, su, fr, rs, ps) in enumerate(draws):
            got = lik.value_grad(bl, su if model else None, fr, rs, ps)
            want = O.loglik_grad(d["peel"], d["tipmask"], d["weights"], model, bl, su, fr, rs, ps, rooted=rooted, dp_eigen=True)
            flat = np.concatenate([[got.log_P], got.grad_blens, got.grad_subst, got.grad_freqs, got.grad_rs, got.grad_ps])
            w = want.flat()
            err = np.abs(flat - w) / np.maximum(1.0, np.abs(w))
            k = int(np.argmax(err))
            print(name, model, C, i, "logL rel %.2e" % (abs(got.log_P - want.logp) / abs(want.logp)), "max grad err %.2e at %d (got %.6g want %.6g)" % (err[1:].max(), k, flat[k], w[k]), "finite", np.all(np.isfinite(flat)))
