import sys, os; sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
from phylostan_b200 import advi, likelihood as lk
d=np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "DS1.npz"))
with lk.TreeLikelihood(d["peel"], d["tipmask"], d["weights"], model="GTR", categories=4, rooted=False) as lik:
    m=advi.UnrootedModel(lik,"GTR")
    rng=np.random.default_rng(1)
    mu=rng.uniform(-2,2,m.dim)
    for t in range(200):
        Z=mu+rng.standard_normal((4,m.dim))
        lp,G=m.log_prob_grad(Z)
        if not (np.all(np.isfinite(lp)) and np.all(np.isfinite(G))):
            c=m.constrain(Z)
            bad=np.argwhere(~np.isfinite(G))
            print("bad", t, lp, bad[:10], len(bad))
            b=bad[0][0] if len(bad) else int(np.argmin(lp))
            rs,_=advi.weibull_rates(c["wshape"][b:b+1],4)
            print("wshape",c["wshape"][b],"rs",rs,"rates",c["rates"][b],"freqs",c["freqs"][b],"blens max/min",c["blens"][b].max(),c["blens"][b].min())
            vg=lik.value_grad(c["blens"][b],c["rates"][b],c["freqs"][b],rs[0],np.full(4,.25))
            print(vg.log_P, "gb nonfinite", np.argwhere(~np.isfinite(vg.grad_blens)).ravel()[:10], vg.grad_subst, vg.grad_freqs, vg.grad_rs, vg.grad_ps)
            np.savez("gpurun_out/bad_draw.npz", blens=c["blens"][b], rates=c["rates"][b], freqs=c["freqs"][b], rs=rs[0])
            break
    else: print("all finite")
