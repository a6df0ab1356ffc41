"""Host-side logic of the product, CPU only: encoders vs the reference's own encoders (golden
.npz), the traversal plan, the per-draw model algebra, and that libphylo_b200.so loads and exports
every symbol include/phylo_b200.h declares (no compute calls without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

from phylostan_b200 import encode as E
from phylostan_b200 import likelihood as lk
from phylostan_b200 import synth

from conftest import ROOT

REF = "/root/reference"


# ------------------------------------------------------------------------------- C ABI surface

def test_library_loads_and_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "phylo_b200.h")).read()
    declared = set(re.findall(r"PHYLO_B200_API [\w\s\*]*?\b(phylo_b200_\w+)\(", header))
    assert len(declared) >= 20
    L = lk.lib()
    for name in sorted(declared):
        assert hasattr(L, name), name
    assert L.phylo_b200_abi_version() == 1


def test_create_fails_loudly_without_a_gpu():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    with pytest.raises(lk.PhyloB200Error) as ei:
        lk.TreeLikelihood(np.array([[1, 2, 3]]), np.array([[1], [2]], dtype=np.uint8), model="JC69")
    assert ei.value.code in (lk.ENODEV, lk.ECUDA)
    assert "no CPU fallback" in str(ei.value) or "CUDA" in str(ei.value)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under phylostan_b200/ may mention it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "phylostan_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", ".sh")):
                src = open(os.path.join(dirpath, f)).read().lower()
                assert "oracle" not in src, (dirpath, f)


# ------------------------------------------------------------------------------- encoders

@pytest.mark.parametrize("name,tf,af,rooted", [
    ("fluA", "examples/fluA/fluA.tree", "examples/fluA/fluA.fa", True),
    ("DS1", "examples/DS1/DS1.trees", "examples/DS1/DS1.nex", False),
    ("HCV", "examples/HCV/HCV.tree", "examples/HCV/HCV.nexus", True)])
def test_encoders_match_reference_utils(datasets, name, tf, af, rooted):
    """phylostan/utils.py:59-104,156-190 run by tests/golden/make_golden_datasets.py vs encode.py."""
    if not os.path.isdir(REF):
        pytest.skip("reference tree not mounted (GPU box)")
    tree = E.read_tree(os.path.join(REF, tf))
    data = E.encode(tree, E.read_alignment(os.path.join(REF, af)), rooted=rooted)
    g = datasets[name]
    np.testing.assert_array_equal(data.peel, g["peel"])
    np.testing.assert_array_equal(data.map, g["map"])
    np.testing.assert_array_equal(data.tipmask, g["tipmask"])
    np.testing.assert_array_equal(data.weights, g["weights"])
    assert list(g["taxa"]) == data.taxa


def test_dataset_shapes_match_survey(datasets):
    """fluA 69x987 -> 238 patterns, DS1 27x1949 -> 934, HCV 63x411 -> 246."""
    for name, S, L, sites in (("fluA", 69, 238, 987), ("DS1", 27, 934, 1949), ("HCV", 63, 246, 411)):
        g = datasets[name]
        assert g["tipmask"].shape == (S, L) and g["peel"].shape == (S - 1, 3)
        assert g["weights"].sum() == sites
        assert g["peel"][-1, 2] == 2 * S - 1
        assert set(np.unique(g["tipmask"])) <= {1, 2, 4, 8, 15}
    assert datasets["DS1"]["peel"][-1, 1] == 2 * 27 - 2  # phylostan.py:264-267


def test_newick_indexing_convention():
    """utils.py:59-81 on eigen/example.tree: tips in order of appearance, internals in post-order."""
    tree = E.parse_newick("((1:0.1,2:0.1):0.2,3:0.3);")
    E.setup_indexes(tree)
    np.testing.assert_array_equal(E.get_peeling_order(tree), [[1, 2, 4], [4, 3, 5]])
    np.testing.assert_array_equal(E.get_preorder(tree), [[5, 0], [4, 5], [1, 4], [2, 4], [3, 5]])
    np.testing.assert_allclose(E.branch_lengths(tree), [0.1, 0.1, 0.3, 0.2])


def test_pattern_compression_and_ambiguity():
    seqs = {"a": "ACGTAC-N", "b": "ACGTACRT", "c": "AAGTAAYT"}
    tipmask, w = E.compress_patterns(seqs, ["a", "b", "c"])
    # columns: AAA CCA GGG TTT AAA CCA -RY NTT -> patterns AAA(2) CCA(2) GGG TTT -RY NTT
    np.testing.assert_array_equal(w, [2, 2, 1, 1, 1, 1])
    np.testing.assert_array_equal(tipmask[0], [1, 2, 4, 8, 15, 15])
    np.testing.assert_array_equal(tipmask[1], [1, 2, 4, 8, 15, 8])   # R is not resolved: utils.py:180-188
    td = E.mask_to_tipdata(tipmask)
    np.testing.assert_array_equal(E.tipdata_to_mask(td), tipmask)
    assert td[0, 4].tolist() == [1, 1, 1, 1]


def test_weibull_rates_mean_one():
    for shape in (0.3, 0.488, 1.0, 3.0):
        for C in (2, 4, 8):
            rs = E.weibull_rates(shape, C)
            assert rs.mean() == pytest.approx(1.0, rel=1e-14) and np.all(np.diff(rs) > 0)


# ------------------------------------------------------------------------------- traversal plan

def _check_plan(peel):
    """Replay the plan on an abstract stack machine: TOS register + shared-memory slots."""
    S = peel.shape[0] + 1
    p = lk.plan(peel)
    post, pre = p["post"], p["pre"]
    TIP, TOS = -1, -2
    children = {int(r[2]) - 1: {int(r[0]) - 1, int(r[1]) - 1} for r in peel}
    # post-order: every internal node once, children first; operands come from where they really are
    done, row_of, slots, tos = set(range(S)), {}, {}, None
    for i, (a, b, sa, sb, spill, node, _, _) in enumerate(post):
        a, b, sa, sb, spill, node = map(int, (a, b, sa, sb, spill, node))
        assert {a, b} == children[node]
        consumed_tos = False
        for ch, src in ((a, sa), (b, sb)):
            assert ch in done
            if ch < S:
                assert src == TIP
            elif src == TOS:
                assert tos == ch, "the TOS must hold this child"
                consumed_tos = True
            else:
                assert slots.pop(src) == ch and src == len(slots), "slots are used as a stack"
        if spill >= 0:
            assert not consumed_tos and tos is not None and spill == len(slots)
            slots[spill] = tos
        else:
            assert consumed_tos or tos is None, "a live TOS may only be overwritten after a spill"
        assert len(slots) <= p["depth_post"]
        tos = node
        done.add(node)
        row_of[node] = i
    assert len(done) == 2 * S - 1 and not slots and tos == 2 * S - 2
    # pre-order: q(node) is where the plan says; q(a) -> TOS, q(b) -> slot
    qslots, tosq, seen = {}, 2 * S - 2, set()
    for node, a, b, sn, db, a_int, rown, rowa, rowb, *_ in pre:
        node, a, b, sn, db, a_int, rown, rowa, rowb = map(int, (node, a, b, sn, db, a_int, rown, rowa, rowb))
        assert {a, b} == children[node] and rown == row_of[node]
        if sn == TOS:
            assert tosq == node
        else:
            assert qslots.pop(sn) == node and sn == len(qslots)
        assert (rowa == row_of[a]) if a >= S else (rowa == -1)
        assert (rowb == row_of[b]) if b >= S else (rowb == -1)
        assert a_int == (1 if a >= S else 0)
        if b >= S:
            assert a >= S and db == len(qslots)      # b internal implies a internal (a descends first)
            qslots[db] = b
        else:
            assert db == -1
        assert len(qslots) <= p["depth_pre"]
        tosq = a if a >= S else None
        seen.add(node)
    assert not qslots and len(seen) == S - 1
    return p


def test_plan_on_reference_trees(datasets):
    for name in ("fluA", "DS1", "HCV"):
        p = _check_plan(datasets[name]["peel"])
        S = datasets[name]["peel"].shape[0] + 1
        assert max(p["depth_post"], p["depth_pre"]) <= int(np.log2(S))


def test_plan_depth_bounds():
    rng = np.random.default_rng(0)
    for S in (2, 3, 5, 64, 257, 1000):
        p = _check_plan(synth.coalescent_peel(S, rng))
        assert max(p["depth_post"], p["depth_pre"]) <= max(int(np.log2(S)), 0)   # Strahler bound minus the TOS
    # caterpillar: depth 1; perfectly balanced 64 tips: depth log2(64) = 6
    S = 50
    cat = [[1, 2, S + 1]] + [[S + k, k + 2, S + k + 1] for k in range(1, S - 1)]
    p = _check_plan(np.array(cat, dtype=np.int32))
    assert p["depth_post"] == 0 and p["depth_pre"] == 0      # a caterpillar never leaves the registers
    rows, nxt, level = [], 65, list(range(1, 65))
    while len(level) > 1:
        new = []
        for i in range(0, len(level), 2):
            rows.append([level[i], level[i + 1], nxt]); new.append(nxt); nxt += 1
        level = new
    # renumber is unnecessary for the plan: children precede parents
    p = _check_plan(np.array(rows, dtype=np.int32))
    assert p["depth_post"] == 5 and p["depth_pre"] == 5      # log2(64) live vectors, one of them in registers


def test_plan_with_table_nodes_as_leaves(datasets):
    """The second plan (message tables in the post-order): table nodes are the internal nodes with at most three tips
    below them; the post-order has no step for them or for anything below them and sees them as tip-like children; the
    pre-order still visits every internal node, knows which ones have no post-order row (no rescale exponents) and gives
    every internal node a parking row of its own."""
    rng = np.random.default_rng(3)
    peels = [datasets[n]["peel"] for n in ("fluA", "DS1", "HCV")] + [synth.coalescent_peel(S, rng) for S in (8, 50, 333)]
    for peel in peels:
        S = peel.shape[0] + 1
        full = lk.plan(peel)
        for max_tips in (2, 3):
            t = lk.plan_tables(peel, max_tips)
            # tips below every node
            ntips = np.ones(2 * S - 1, dtype=int)
            for a, b, p in peel:
                ntips[p - 1] = ntips[a - 1] + ntips[b - 1]
            is_tab = t["node_tab"] >= 0
            assert np.array_equal(is_tab, (np.arange(2 * S - 1) >= S) & (ntips <= max_tips))
            assert t["table_nodes"] == int(is_tab.sum())
            assert t["table_entries"] == 25 * int((is_tab & (ntips == 2)).sum()) + 125 * int((is_tab & (ntips == 3)).sum())
            # post-order: one step per internal node that is not a table node, children before parents, leafified
            # children tip-like
            post = t["post"]
            assert t["post_steps"] == post.shape[0] == S - 1 - t["table_nodes"]
            done = set(range(S)) | set(np.nonzero(is_tab)[0].tolist())
            for a, b, sa, sb, spill, node, _, _ in post:
                assert a in done and b in done and not is_tab[node]
                if is_tab[a] or a < S:
                    assert sa == -1
                if is_tab[b] or b < S:
                    assert sb == -1
                done.add(int(node))
            assert int(post[-1][5]) == 2 * S - 2 and t["depth"] <= max(full["depth_post"], full["depth_pre"])
            # pre-order: the same traversal as the full plan; rows only where the post-order has a step; parking rows
            pre = t["pre"]
            assert np.array_equal(pre[:, :6], full["pre"][:, :6])
            row_of = {int(r[5]): i for i, r in enumerate(post)}
            parks = set()
            for node, a, b, src, dst_b, a_int, rown, rowa, rowb, parkn, parkb, _ in pre:
                assert rown == row_of.get(int(node), -1)
                assert rowa == row_of.get(int(a), -1) and rowb == row_of.get(int(b), -1)
                assert 0 <= parkn < S - 1 and (parkn == rown or rown < 0)
                parks.add(int(parkn))
                if b >= S:
                    assert 0 <= parkb < S - 1
            assert len(parks) == S - 1   # one parking row per internal node
    with pytest.raises(lk.PhyloB200Error):
        lk.plan_tables(np.array([[1, 2, 4], [4, 3, 5]], dtype=np.int32))   # three taxa: the root is a pitchfork


def test_plan_rejects_malformed_peel():
    with pytest.raises(lk.PhyloB200Error):
        lk.plan(np.array([[4, 3, 5], [1, 2, 4]], dtype=np.int32))      # parent before child
    with pytest.raises(lk.PhyloB200Error):
        lk.plan(np.array([[1, 2, 4], [1, 3, 5]], dtype=np.int32))      # node used twice
    with pytest.raises(lk.PhyloB200Error):
        lk.plan(np.array([[1, 2, 5], [5, 3, 4]], dtype=np.int32))      # root is not 2S-1


# ------------------------------------------------------------------------------- model algebra

def test_derive_reconstructs_q_and_stationarity():
    rng = np.random.default_rng(2)
    for model, subst in (("JC69", None), ("HKY", [5.58]), ("GTR", rng.dirichlet(np.ones(6)))):
        fr = None if model == "JC69" else rng.dirichlet(np.ones(4) * 4)
        d = lk.derive(model, subst, fr)
        Q, pi = d["Q"], d["pi"]
        np.testing.assert_allclose(Q.sum(1), 0, atol=1e-15)
        assert -(np.diag(Q) * pi).sum() == pytest.approx(1.0, rel=1e-14)          # generate_script.py:868
        np.testing.assert_allclose(d["m1"] @ np.diag(d["lam"]) @ d["m2"], Q, atol=1e-14)
        np.testing.assert_allclose(d["m1"] @ d["m2"], np.eye(4), atol=1e-14)
        np.testing.assert_allclose(pi @ Q, 0, atol=1e-15)
        assert np.all(np.diff(d["lam"]) >= 0) and abs(d["lam"][-1]) < 1e-14           # eigenvalues_sym order
    un = lk.derive("JC69", normalize=False)["Q"]                                      # eigen/eigen.py:47-50
    np.testing.assert_allclose(un, np.full((4, 4), 0.25) - np.eye(4), atol=1e-16)


def test_derive_x_theta_matches_finite_differences():
    """X_theta = m2 dQ/dtheta m1 for every unconstrained parameter (rates, kappa, freqs)."""
    rng = np.random.default_rng(4)
    for model, subst in (("HKY", np.array([3.3])), ("GTR", rng.dirichlet(np.ones(6)) * 6)):
        fr = rng.dirichlet(np.ones(4) * 4)
        d = lk.derive(model, subst, fr)
        theta = np.concatenate([subst, fr])
        for k in range(theta.size):
            h = 1e-6
            tp, tm = theta.copy(), theta.copy()
            tp[k] += h; tm[k] -= h
            n = subst.size
            dQ = (lk.derive(model, tp[:n], tp[n:])["Q"] - lk.derive(model, tm[:n], tm[n:])["Q"]) / (2 * h)
            np.testing.assert_allclose(d["m1"] @ d["X"][k] @ d["m2"], dQ, atol=2e-9)


def test_message_statistic_contraction_identities():
    """The algebra behind the message-statistic sweep, on the library's own eigen-system:
    with G~ = G P^T,  <G, Q P> = <G~, Q>  and  <G, dP/dtheta> = <m1^T G~ m2^T o Phi, X_theta>,
    Phi_ij = (e^{(l_i - l_j) tau} - 1) / (l_i - l_j), Phi_ii = tau -- and the rounding amplification e^{(l_i - l_j) tau}
    that makes the library fall back to the plain statistic on long branches."""
    from scipy.linalg import expm
    rng = np.random.default_rng(8)
    for model, subst in (("HKY", np.array([4.2])), ("GTR", rng.dirichlet(np.ones(6)) * 6)):
        fr = rng.dirichlet(np.ones(4) * 4)
        d = lk.derive(model, subst, fr)
        Q, lam, m1, m2 = d["Q"], d["lam"], d["m1"], d["m2"]
        theta = np.concatenate([subst, fr])
        for tau in (1e-6, 0.03, 0.7, 4.0):
            P = expm(Q * tau)
            G = rng.normal(size=(4, 4))
            Gt = G @ P.T
            assert np.sum(G * (Q @ P)) == pytest.approx(np.sum(Gt * Q), rel=1e-11, abs=1e-13)
            x = (lam[:, None] - lam[None, :]) * tau
            with np.errstate(invalid="ignore", divide="ignore"):
                Phi = tau * np.where(np.abs(x) < 1e-8, 1.0 + 0.5 * x, np.expm1(x) / x)
            Ht = m1.T @ Gt @ m2.T
            for k in range(theta.size):
                h = 1e-6
                tp, tm = theta.copy(), theta.copy()
                tp[k] += h; tm[k] -= h
                n = subst.size
                dP = (expm(lk.derive(model, tp[:n], tp[n:])["Q"] * tau) - expm(lk.derive(model, tm[:n], tm[n:])["Q"] * tau)) / (2 * h)
                want = np.sum(G * dP)
                got = np.sum(Ht * Phi * d["X"][k])
                assert got == pytest.approx(want, rel=2e-7, abs=2e-8)
        # where it stops being usable: e^{spread * tau} of rounding amplification
        tau = 40.0 / np.ptp(lam)
        assert np.exp(np.ptp(lam) * tau) * 2.0 ** -53 > 1e-2


def test_derive_rejects_out_of_domain():
    with pytest.raises(lk.PhyloDomainError):
        lk.derive("HKY", [float("nan")], [0.25] * 4)
    with pytest.raises(lk.PhyloDomainError):
        lk.derive("GTR", np.ones(6), [0.5, 0.5, 0.0, 0.0])


def test_ratio_transform_helpers_match_the_stan_loops():
    """phylo_b200_ratios_forward / _reverse (host only): generate_script.py:711-752 restated in Python, the
    reverse sweep against finite differences of  sum(c * heights) + log-Jacobian."""
    d = np.load(os.path.join(ROOT, "tests", "golden", "fluA.npz"))
    S = d["tipmask"].shape[0]
    nn = 2 * S - 1
    rng = np.random.default_rng(4)
    lowers = np.zeros(nn)
    lowers[:S] = rng.uniform(0, 5, S)
    for node, par in d["map"][:0:-1]:
        lowers[int(par) - 1] = max(lowers[int(par) - 1], lowers[int(node) - 1])
    B = 3
    props, root = rng.uniform(0.05, 0.95, (B, S - 2)), lowers.max() + rng.uniform(1, 5, B)
    h, lj = lk.ratios_forward(d["map"], lowers, props, root)
    for b in range(B):                                       # the Stan function `transform` + the Jacobian loop
        hs, j, logj = np.zeros(S - 1), 0, 0.0
        hs[int(d["map"][0, 0]) - S - 1] = root[b]
        for node, par in d["map"][1:]:
            if node > S:
                lo = lowers[node - 1]
                hs[node - S - 1] = lo + (hs[par - S - 1] - lo) * props[b, j]
                logj += np.log(hs[par - S - 1] - lo)
                j += 1
        assert np.allclose(h[b], hs, rtol=1e-14) and lj[b] == pytest.approx(logj, rel=1e-13)
    c = rng.normal(0, 1, S - 1)
    f = lambda p, r: (lk.ratios_forward(d["map"], lowers, p, r)[0] * c).sum(axis=1) + lk.ratios_forward(d["map"], lowers, p, r)[1]
    gp, gr = lk.ratios_reverse(d["map"], lowers, props, h, np.tile(c, (B, 1)))
    for k in (0, 7, S - 3):
        e = np.zeros((B, S - 2)); e[:, k] = 1e-5
        assert np.allclose((f(props + e, root) - f(props - e, root)) / 2e-5, gp[:, k], rtol=1e-5, atol=1e-6)
    assert np.allclose((f(props, root + 1e-4) - f(props, root - 1e-4)) / 2e-4, gr, rtol=1e-5, atol=1e-6)
    # contemporaneous tips: lowers = NULL
    h0, _ = lk.ratios_forward(d["map"], None, props, root)
    assert np.all(h0 > 0) and np.all(h0 <= root[:, None] + 1e-12)
    bad = np.array(d["map"]).copy()
    bad[5, 1] = 1                                            # a tip as parent
    with pytest.raises(lk.PhyloB200Error):
        lk.ratios_forward(bad, lowers, props, root)


def test_front_end_map_validation_is_shared_by_forward_and_reverse():
    """One validation for every front-end entry (round-1 advice: ratios_reverse indexed with unchecked map rows, and a
    duplicated or missing node went unnoticed): row 0 = root, every node once, after its parent, two children each."""
    d = np.load(os.path.join(ROOT, "tests", "golden", "fluA.npz"))
    S = d["tipmask"].shape[0]
    rng = np.random.default_rng(8)
    props, root = rng.uniform(0.1, 0.9, (2, S - 2)), np.array([30.0, 40.0])
    h, _ = lk.ratios_forward(d["map"], None, props, root)
    good = np.array(d["map"])

    def broken(edit):
        m = good.copy()
        edit(m)
        return m
    cases = {
        "root not in row 0": broken(lambda m: m.__setitem__((0, 0), int(m[1, 0]))),
        "node twice": broken(lambda m: m.__setitem__(3, m[2].copy())),
        "child before its parent": broken(lambda m: m.__setitem__(slice(1, None), m[:0:-1].copy())),
        "out of range": broken(lambda m: m.__setitem__((4, 0), 2 * S + 5)),
        "tip as parent": broken(lambda m: m.__setitem__((5, 1), 1)),
    }
    for name, m in cases.items():
        with pytest.raises(lk.PhyloB200Error):
            lk.ratios_forward(m, None, props, root)
        with pytest.raises(lk.PhyloB200Error):
            lk.ratios_reverse(m, None, props, h, np.ones((2, S - 1)))
