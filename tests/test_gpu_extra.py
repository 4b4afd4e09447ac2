"""-m gpu: cases the round-1 review found untested — the largest range the reference allows (S = FAMILYSIZEMAX + 1 = 1001),
K1 over every distinct key of the BASELINE configs[2] tree, re-use of one context for a larger tree and range
(buffer re-sizing), and the `pvalue -o` / `pvalue -i` round trip (cafe/pvalue.cpp:63-93)."""
import os

import numpy as np
import pytest

import oracle
from cafe_b200 import gpu as cgpu
from cafe_b200 import host as chost
from util import EXAMPLE_TREE, Problem, rel_err

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_largest_range_S_1001():
    # max family size 800 -> root range 1..1000 (rint(1.25 * 800)), family range 0..960: S = 1001 = FAMILYSIZEMAX + 1.
    # All branches have the same length, so the CPU oracle builds ONE 1001 x 1001 matrix (~30 s) for the whole test.
    rng = np.random.RandomState(5)
    base = rng.randint(700, 790, size=(24, 1))
    counts = np.clip(base + rng.randint(-12, 13, size=(24, 4)), 0, 800).astype(np.int32)
    counts[0, 0] = 800
    p = Problem("((a:40,b:40):40,(c:40,d:40):40)", counts, 0.0015, prior_lambda=750.0)
    assert p.ranges == (0, 960, 1, 1000) and p.maxfs == 1000
    g = p.make_gpu()
    mats = p.oracle_mats()                   # one distinct key: one CPU matrix
    ref = mats[0]
    o = oracle.score(p.otree, mats, p.counts, p.ranges, p.prior, want_L=True)
    for node in (0, 4):
        M = g.get_matrix(node)
        assert M.shape == (1001, 1001)
        big = ref > 1e-280
        assert rel_err(M[big], ref[big]).max() < 1e-12
    score, fz = g.score()
    L = g.family_likelihoods()
    assert fz == o["first_zero"]
    big = o["L"] > 1e-290
    assert big.any() and rel_err(L[big], o["L"][big]).max() < 1e-11
    if np.isfinite(o["score"]):
        assert abs(score - o["score"]) <= max(1e-6, 1e-12 * abs(o["score"]))
    else:
        assert score == o["score"]
    g.close()


def test_k1_every_key_of_the_config2_tree():
    import bench_data
    nw = bench_data.config_tree("configs[2]")
    lam0 = bench_data.default_lambda(nw)
    tree = chost.parse_tree(nw)
    rg = chost.init_family_size(400)
    ranges = (rg["min"], rg["max"], rg["root_min"], rg["root_max"])
    maxfs = max(ranges[1], ranges[3])
    n = tree.n_nodes
    g = cgpu.CafeGpu()
    g.set_tree(tree.left, tree.right, tree.branchlength)
    g.set_ranges(*ranges)
    g.set_lnc_table(chost.lnc_table(maxfs))
    g.set_rates(np.full(n, lam0), np.full(n, 0.8 * lam0))
    g.build_matrices()
    seen = {}
    for v in range(n):
        if v == tree.root:
            continue
        seen.setdefault(int(tree.branchlength[v]), v)
    assert len(seen) == g.num_keys() >= 40
    # every key: the first row (exact, birthdeath.c:244) and the structure of a transition matrix (entries in [0, 1], rows of
    # small sizes sum to 1 within the truncation); every fifth key, the shortest and the longest branch: every entry against the
    # CPU oracle (a 501 x 501 matrix takes the oracle ~4 s)
    keys = sorted(seen.items())
    exact = set(keys[::5]) | {keys[0], keys[-1]}
    worst = 0.0
    for t, v in keys:
        M = g.get_matrix(v)
        assert M.min() >= 0.0 and M.max() <= 1.0 and M[0, 0] == 1.0 and (M[0, 1:] == 0).all()
        assert np.abs(M[1:40].sum(axis=1) - 1.0).max() < 1e-9
        if (t, v) in exact:
            ref = oracle.bd_matrix(t, lam0, 0.8 * lam0, maxfs)
            big = ref > 1e-280
            worst = max(worst, rel_err(M[big], ref[big]).max())
            assert np.array_equal(M[0], ref[0])
    assert worst < 1e-12, worst
    g.close()


def test_one_context_reused_for_a_larger_tree_and_range():
    rng = np.random.RandomState(2)
    small = Problem("((a:3,b:3):4,c:7)", np.maximum(0, 6 + rng.randint(-3, 4, size=(128, 3))).astype(np.int32), 0.01)
    nw_big = oracle.random_tree(9, 3)
    base = rng.randint(20, 60, size=(128, 1))
    big = Problem(nw_big, np.maximum(0, base + rng.randint(-5, 6, size=(128, 9))).astype(np.int32), 0.004)
    assert big.ranges[1] > small.ranges[1] and big.ranges[3] > small.ranges[3]
    g = small.make_gpu()
    s_small, _ = g.score()
    assert abs(s_small - small.oracle_score(want_L=False)["score"]) < 1e-6
    # the same context, now with more leaves, a wider range and the same number of families (F_pad unchanged)
    t = big.tree
    g.set_tree(t.left, t.right, t.branchlength)
    g.set_ranges(*big.ranges)
    with pytest.raises(cgpu.CafeGpuError):
        g.set_rates(big.lam_node, big.mu_node); g.build_matrices(); g.score()   # families and prior of the old geometry are gone
    g.set_lnc_table(chost.lnc_table(big.maxfs))
    g.set_families(big.counts)
    g.set_prior(big.prior)
    E = np.eye(big.ranges[1] + 1)
    g.set_error_model(8, E)      # an identity error model on the last leaf: exercises the per-leaf matrices behind the keys
    g.set_rates(big.lam_node, big.mu_node)
    g.build_matrices()
    s_big, fz = g.score()
    fresh = big.make_gpu()
    s_fresh, _ = fresh.score()
    assert fz == -1 and s_big == s_fresh
    assert abs(s_big - big.oracle_score(want_L=False)["score"]) <= max(1e-6, 1e-12 * abs(s_big))
    # and back to the small problem
    g.set_tree(small.tree.left, small.tree.right, small.tree.branchlength)
    g.set_ranges(*small.ranges)
    g.set_lnc_table(chost.lnc_table(small.maxfs))
    g.set_families(small.counts)
    g.set_prior(small.prior)
    g.set_rates(small.lam_node, small.mu_node)
    g.build_matrices()
    assert g.score()[0] == s_small
    g.close()
    fresh.close()


def test_pvalue_out_in_round_trip(tmp_path):
    z = np.load(os.path.join(GOLD, "example.npz"))
    species = [str(s) for s in z["species_leaf_order"]]
    tab = str(tmp_path / "example_data.tab")
    with open(tab, "w") as f:
        f.write("\t".join(["FAMILYDESC", "FAMILY"] + species) + "\n")
        for i, r in zip(z["ids"], z["counts"]):
            f.write("\t".join(["d", str(i)] + [str(x) for x in r]) + "\n")
    s = chost.Session(quiet=True)
    assert s.command("seed 10") == 0
    assert s.command("load -i %s -t 1 -r 60" % tab) == 0
    assert s.command("tree " + EXAMPLE_TREE) == 0
    assert s.command("lambda -l 0.005") == 0
    out = str(tmp_path / "cd.txt")
    assert s.command("pvalue -o " + out) == 0
    cd = s.cond_dist()
    rg = s.ranges()
    assert cd.shape == (rg["root_max"] - rg["root_min"] + 1, 60)
    assert (np.diff(cd, axis=1) >= 0).all() and cd.min() >= 0 and cd.max() <= 1     # rows ascending probabilities
    lines = open(out).read().splitlines()
    assert len(lines) == cd.shape[0] and all(len(ln.split("\t")) == 60 for ln in lines)
    # a second session reads the file back: the matrix it then uses equals the written one to the 9 printed digits
    s2 = chost.Session(quiet=True)
    assert s2.command("load -i %s -t 1 -r 60" % tab) == 0
    assert s2.command("tree " + EXAMPLE_TREE) == 0
    assert s2.command("pvalue -i " + out) == 0
    cd2 = s2.cond_dist()
    assert cd2.shape == cd.shape
    np.testing.assert_allclose(cd2, cd, rtol=1e-8, atol=1e-300)
    s.close()
    s2.close()


def test_k1_recurrence_equals_the_term_by_term_kernel(monkeypatch):
    # K1's default kernel carries 15 of every 16 terms by the ratio of consecutive terms (bd_matrix.cu); CAFE_GPU_K1_EXACT=1
    # evaluates every term with its own exp(), the reference's arithmetic operation for operation.  They must agree within the
    # 1e-12 the matrices are held to (measured on the bench shapes, tools/k1_check.py: <= 3.7e-13 up to S = 1001), entries
    # below the double range included (no term lost to an underflowing anchor), and a key outside the recurrence's guard
    # (lambda t ~ 1e-7: q = coeff / (alpha beta) ~ 1e14) must be bit-identical because it takes the term-by-term path.
    newick = "(((a:31,b:31):12,c:43):20,(d:7,(e:2,f:2):5):56)"
    tree = chost.parse_tree(newick)
    n = tree.n_nodes
    for maxsize, lam, mu in ((120, 0.004, None), (260, 0.0021, 0.0017), (260, 0.0005, 0.0005), (120, 1e-8, None)):
        rg = chost.init_family_size(maxsize)
        ranges = (rg["min"], rg["max"], rg["root_min"], rg["root_max"])
        maxfs = max(ranges[1], ranges[3])
        mats = {}
        for mode in ("rec", "exact"):
            if mode == "exact":
                monkeypatch.setenv("CAFE_GPU_K1_EXACT", "1")
            g = cgpu.CafeGpu()
            g.set_tree(tree.left, tree.right, tree.branchlength)
            g.set_ranges(*ranges)
            g.set_lnc_table(chost.lnc_table(maxfs))
            g.set_rates(np.full(n, lam), np.full(n, -1.0 if mu is None else mu))
            g.build_matrices()
            mats[mode] = [g.get_matrix(v) for v in range(n) if v != tree.root]
            g.close()
            monkeypatch.delenv("CAFE_GPU_K1_EXACT", raising=False)
        for a, b in zip(mats["rec"], mats["exact"]):
            if lam == 1e-8:
                assert np.array_equal(a, b)
                continue
            big = b > 1e-300
            assert rel_err(a[big], b[big]).max() < 1e-12
            assert np.abs(a - b)[~big].max() < 1e-299
            assert np.array_equal(a[0], b[0])


def test_k1_recurrence_fuzz_over_rates_and_branch_lengths(monkeypatch):
    # 96 random keys (t, lambda, mu) from lambda t ~ 1e-7 (outside the recurrence's guard) to lambda t ~ 30 (alpha -> 1, coeff <= 0:
    # the zero matrix), mu from 0 (NaN terms, clamped like the reference's macros) to 5 lambda, and mu < 0: the default kernel
    # against the term-by-term one on a 48-leaf caterpillar whose every branch has its own rates.
    rng = np.random.RandomState(11)
    n_leaves = 48
    nw = "L0:1"
    for k in range(1, n_leaves):
        nw = f"({nw},L{k}:1):1"
    nw = nw[: nw.rfind(":")]
    tree = chost.parse_tree(nw)
    n = tree.n_nodes
    bl = np.array(tree.branchlength, dtype=np.float64)
    nonroot = [v for v in range(n) if v != tree.root]
    bl[nonroot] = np.round(10 ** rng.uniform(0, 3, size=len(nonroot)))            # 1 .. 1000
    lam = np.full(n, 0.001); mu = np.full(n, -1.0)
    lt = 10 ** rng.uniform(-7, 1.5, size=len(nonroot))
    lam[nonroot] = lt / bl[nonroot]
    kind = rng.randint(0, 4, size=len(nonroot))                                     # 0: mu < 0, 1: mu = lambda * u, 2: mu = 0, 3: mu = lambda
    for i, v in enumerate(nonroot):
        mu[v] = (-1.0, lam[v] * 10 ** rng.uniform(-2, 0.7), 0.0, lam[v])[kind[i]]
    ranges = (0, 140, 1, 150)
    maxfs = 150
    mats = {}
    for mode in ("rec", "exact"):
        if mode == "exact":
            monkeypatch.setenv("CAFE_GPU_K1_EXACT", "1")
        g = cgpu.CafeGpu()
        g.set_tree(tree.left, tree.right, bl)
        g.set_ranges(*ranges)
        g.set_lnc_table(chost.lnc_table(maxfs))
        g.set_rates(lam, mu)
        g.build_matrices()
        mats[mode] = [g.get_matrix(v) for v in nonroot]
        g.close()
        monkeypatch.delenv("CAFE_GPU_K1_EXACT", raising=False)
    worst = 0.0
    for a, b in zip(mats["rec"], mats["exact"]):
        assert not np.isnan(a).any() and not np.isnan(b).any()
        big = b > 1e-300
        if big.any():
            worst = max(worst, rel_err(a[big], b[big]).max())
        assert np.abs(a - b)[~big].max() < 1e-299
    assert worst < 1e-12, worst
    # and two of the keys against the CPU oracle (the reference's arithmetic), one per summation mode
    for want in (0, 1):
        i = next(i for i in range(len(nonroot)) if kind[i] == want and 1e-3 < lt[i] < 1.0)
        v = nonroot[i]
        ref = oracle.bd_matrix(int(bl[v]), lam[v], mu[v], maxfs)
        big = ref > 1e-280
        assert rel_err(mats["rec"][i][big], ref[big]).max() < 1e-12
