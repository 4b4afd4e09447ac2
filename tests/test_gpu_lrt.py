"""-m gpu: branch-stretch likelihood-ratio test (cafe_gpu_likelihood_ratio_test) against the CPU oracle, which
tests/test_oracle.py pins bit for bit against the compiled reference (cafe/cafe_main.c:342-431), and against the committed output of
the reference itself on example_data.tab (tests/golden/lrt.npz).

Tolerances: the best / base maximum likelihoods within 1e-11 relative (fp64 GEMM summation order, CUDA exp/log); the number of
lengthenings is an integer decision `prev < next` and must agree except where two successive likelihoods tie to within that
rounding (at most 1 % of the (branch, family) pairs); the likelihood ratio 1 - chi2cdf(2 ln(best/base)) within 1e-8 absolute where
the step counts agree."""
import os

import numpy as np
import pytest

import oracle
from cafe_b200 import host as chost

from util import EXAMPLE_TREE, Problem, random_tree, rel_err

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _counts(n_leaves, F, hi, seed, outliers=3):
    rng = np.random.RandomState(seed)
    base = rng.randint(1, hi, size=(F, 1))
    c = np.maximum(0, base + rng.randint(-3, 4, size=(F, n_leaves))).astype(np.int32)
    for k in range(outliers):          # clear outliers: several lengthenings on one branch
        c[rng.randint(F), rng.randint(n_leaves)] += 15 + 5 * k
    return c


def _ratio(best, base):
    return 1.0 if best == base else 1 - chost.chi2cdf(2 * (np.log(best) - np.log(base)), 1)


def _check(p, tested=None, stock_mu=False):
    g = p.make_gpu()
    mu_len = np.zeros(p.otree.n_nodes) if stock_mu else p.mu_node
    base, best, steps = g.likelihood_ratio_test(tested, mu_len if stock_mu else None)
    score_after = g.score()[0]           # the context is back at the tree's own matrices
    g.close()
    assert np.isclose(score_after, p.oracle_score(want_L=False)["score"], rtol=1e-12, atol=1e-6)
    mats = p.oracle_mats()
    le = p.oracle_leaf_err()
    t = p.otree
    bl = np.array(t.branchlength, dtype=np.float64)
    if tested is not None and 2 in tested:
        bl = np.floor(bl)        # a later shard of a split table: nobody here is the table's first tested family
    F = len(p.counts)
    n_pairs = n_step_diff = 0
    max_steps = 0
    for f in range(F):
        if tested is not None and not tested[f]:
            assert np.array_equal(best[:, f][np.arange(t.n_nodes) != t.root], np.full(t.n_nodes - 1, base[f]))
            assert not steps[:, f].any()
            continue
        r_o, best_o, steps_o = oracle.lrt_family(t, mats, p.lam_node, mu_len, bl, p.counts[f], p.ranges, leaf_err=le)
        assert best[t.root, f] == -1 and r_o[t.root] == -1
        for b in range(t.n_nodes):
            if b == t.root:
                continue
            n_pairs += 1
            if steps[b, f] != steps_o[b]:
                n_step_diff += 1
                continue
            max_steps = max(max_steps, int(steps_o[b]))
            assert rel_err(best[b, f], best_o[b]) <= 1e-11, (b, f, best[b, f], best_o[b])
            if steps_o[b] == 0:
                assert best[b, f] == base[f]      # the `prev == maxlh` branch must be exact
            assert abs(_ratio(best[b, f], base[f]) - r_o[b]) <= 1e-8
    assert n_step_diff <= max(1, n_pairs // 100), (n_step_diff, n_pairs)
    return max_steps


def test_lrt_example_tree():
    assert _check(Problem(EXAMPLE_TREE, _counts(5, 64, 25, 3), 0.005)) >= 2


def test_lrt_two_classes_lambda_mu():
    # per-node (lambda, mu): the lengthened branch keeps ITS node's rates (the reference's tree copy drops mu, see oracle/ref_shim.cpp)
    _check(Problem(EXAMPLE_TREE, _counts(5, 40, 20, 4), [0.004, 0.007], mu=[0.003, 0.005], lambda_tree="(((2,2)1,(1,1)1)1,1)"))


def test_lrt_stock_reference_mu():
    # the unmodified binary keys lengthened branches (t, lambda, 0): 0/1 matrices out of clamped NaNs; same numbers here
    _check(Problem(EXAMPLE_TREE, _counts(5, 48, 25, 5), 0.005), stock_mu=True)
    _check(Problem(random_tree(9, 2), _counts(9, 30, 20, 6), 0.004, mu=0.003), stock_mu=True)


@pytest.mark.parametrize("n_leaves,seed", [(3, 2), (13, 4), (20, 5)])
def test_lrt_random_trees(n_leaves, seed):
    nw = random_tree(n_leaves, seed)
    _check(Problem(nw, _counts(n_leaves, 40, 30, seed), 0.01))


def test_lrt_fractional_branch_lengths_and_filter():
    # family 0 is filtered, so the parsed (fractional) lengths go to family 1; every later family starts from (int) lengths
    nw = "(((chimp:6.6,human:6.6):81.2,(mouse:17.4,rat:17.4):70.4):6.9,dog:93.7)"
    c = _counts(5, 37, 22, 8)
    tested = np.ones(len(c), dtype=np.uint8)
    tested[[0, 5, 11]] = 0
    _check(Problem(nw, c, 0.006), tested)
    _check(Problem(nw, c, 0.006), tested * 2)      # the same shard when an earlier one owns the first tested family


def test_lrt_root_range_wider_than_vector_and_error_model():
    rg = chost.init_family_size(30)
    dim = rg["max"] + 1
    E = np.zeros((dim, dim))
    for j in range(dim):
        E[j, j] = 0.9
        E[max(j - 1, 0), j] += 0.05
        E[min(j + 1, dim - 1), j] += 0.05
    nw = random_tree(7, 3)
    c = np.minimum(_counts(7, 30, 20, 6, outliers=1), 30)
    _check(Problem(nw, c, 0.008, ranges=(rg["min"], rg["max"], rg["root_min"], rg["root_max"]), err={0: E, 3: E}))
    _check(Problem(random_tree(9, 4), _counts(9, 33, 25, 9), 0.008, ranges=(0, 70, 3, 97)))


def test_lrt_many_families_all_tiles():
    # more families than one CTA tile pair per SM: every CTA of the fused kernel takes part; sample-checked against the oracle
    nw = random_tree(8, 6)
    c = np.maximum(_counts(8, 30000, 30, 12, outliers=4), 1)   # no empty leaves: those lengthen ~130 times (until (int) overflows)
    p = Problem(nw, c, 0.01)
    g = p.make_gpu()
    base, best, steps = g.likelihood_ratio_test()
    g.close()
    t = p.otree
    mats = p.oracle_mats()
    bl = np.floor(np.array(t.branchlength, dtype=np.float64))
    rng = np.random.RandomState(1)
    for f in rng.choice(len(c), 12, replace=False):
        r_o, best_o, steps_o = oracle.lrt_family(t, mats, p.lam_node, p.mu_node, bl, c[f], p.ranges)
        same = steps[:, f] == steps_o
        assert same.mean() >= 0.9
        nz = same & (np.arange(t.n_nodes) != t.root)
        assert rel_err(best[nz, f], best_o[nz]).max() <= 1e-11
    # duplicates of a pattern give identical bits wherever they sit in the table
    c2 = np.concatenate([c[:100], c[:100]])
    p2 = Problem(nw, c2, 0.01, ranges=p.ranges)
    g = p2.make_gpu()
    _, best2, steps2 = g.likelihood_ratio_test()
    g.close()
    assert np.array_equal(best2[:, :100], best2[:, 100:]) and np.array_equal(steps2[:, :100], steps2[:, 100:])
    assert np.array_equal(best2[:, :100], best[:, :100])


def test_lrt_reference_golden_through_the_host_mirror(tmp_path):
    # `report <name> likelihood` (reports.cpp:684-685) on example_data.tab: the reference's own likelihoodRatios, committed by
    # tests/golden/make_golden.py (lambda = 0.005, family p-values of cond_dist.npz, cut-off 0.05)
    z = np.load(os.path.join(GOLD, "example.npz"))
    g = np.load(os.path.join(GOLD, "lrt.npz"))
    species = [str(s) for s in z["species_leaf_order"]]
    path = str(tmp_path / "example_data.tab")
    with open(path, "w") as f:
        f.write("\t".join(["FAMILYDESC", "FAMILY"] + species) + "\n")
        for i, r in zip(z["ids"], z["counts"]):
            f.write("\t".join(["d", str(i)] + [str(x) for x in r]) + "\n")
    s = chost.Session(quiet=True)
    assert s.command("seed 10") == 0
    assert s.command("load -i %s -t 1 -p %g" % (path, float(g["cutoff"]))) == 0
    assert s.command("tree " + EXAMPLE_TREE) == 0
    assert s.command("lambda -l %.10g" % float(g["lam"])) == 0
    s.set_max_pvalues(g["max_pvalues"])
    lr = s.likelihood_ratio_test()
    lr_stock = s.likelihood_ratio_test(tree_level_mu=True)
    # the command itself: `report <name> likelihood` draws its own conditional distribution, filters on ITS family p-values and
    # writes one line per family
    rep = str(tmp_path / "rep")
    assert s.command("report %s likelihood" % rep) == 0
    lr_cmd = s.likelihood_ratio_test()          # same p-values as the report just used
    lines = [ln for ln in open(rep + ".cafe").read().split("\n") if ln.startswith("ENSF")]
    assert [ln.split("\t")[0] for ln in lines] == [str(i) for i in z["ids"]]
    written = np.array([[-1.0 if x == "-" else float(x) for x in ln.split("\t")[4].strip("()").split(",")] for ln in lines]).T
    assert written.shape == lr_cmd.shape and np.allclose(written, lr_cmd, rtol=1e-5, atol=1e-12)
    assert s.command("report %s lh2" % rep) != 0                 # stops inside the reference itself; rejected here
    s.close()
    # "ratios": the reference with its tree copy carrying the nodes' mu (oracle/ref_shim.cpp); "ratios_stock": the unmodified
    # behaviour, the numbers `report <name> likelihood` of the stock binary prints
    for got, ref in ((lr, g["ratios"]), (lr_stock, g["ratios_stock"])):
        assert got.shape == ref.shape
        assert np.array_equal(got == -1, ref == -1)          # root row and filtered families
        assert np.array_equal(got == 1, ref == 1)            # "no lengthening helped"
        assert np.abs(got - ref).max() <= 1e-8
    assert np.abs(lr - lr_stock).max() > 1e-3


def test_lrt_error_states():
    from cafe_b200 import gpu as cgpu
    p = Problem(EXAMPLE_TREE, _counts(5, 16, 20, 2), 0.005)
    g = p.make_gpu()
    g.set_rates(p.lam_node, p.mu_node)                       # new rates, matrices not rebuilt yet
    with pytest.raises(cgpu.CafeGpuError, match="build_matrices"):
        g.likelihood_ratio_test()
    g.build_matrices()
    g.set_key_shard(0, 2)                                    # K1 sharded over two ranks: the test needs all matrices locally
    g.build_matrices()
    with pytest.raises(cgpu.CafeGpuError):
        g.likelihood_ratio_test()
    g.set_key_shard(0, 1)
    g.build_matrices()
    base, best, steps = g.likelihood_ratio_test(np.zeros(16, dtype=np.uint8))   # nobody tested: nothing to do
    root = p.otree.root
    assert not steps.any() and np.array_equal(np.delete(best, root, axis=0), np.tile(base, (p.otree.n_nodes - 1, 1)))
    assert (best[root] == -1).all()
    s_after, _ = g.score()
    assert abs(s_after - p.oracle_score(want_L=False)["score"]) < 1e-6
    g.close()
