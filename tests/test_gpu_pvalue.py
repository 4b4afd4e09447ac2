"""-m gpu: K4 (conditional distribution) and K5 (family p-values) through the C-ABI against the oracle."""
import numpy as np
import pytest

import oracle

from util import EXAMPLE_TREE, Problem, rel_err

pytestmark = pytest.mark.gpu


def small_counts(n_leaves, F, hi, seed):
    rng = np.random.RandomState(seed)
    base = rng.randint(1, hi, size=(F, 1))
    return np.maximum(0, base + rng.randint(-2, 3, size=(F, n_leaves))).astype(np.int32)


def make(newick=EXAMPLE_TREE, lam=0.004, F=24, hi=12, ranges=(0, 62, 1, 15), seed=1, **kw):
    n_leaves = newick.count(",") + 1
    return Problem(newick, small_counts(n_leaves, F, hi, seed), lam, ranges=ranges, **kw)


@pytest.mark.parametrize("newick,lam", [
    (EXAMPLE_TREE, 0.004),
    ("((A:10,B:10):5,(C:7,D:7):8)", 0.01),
    ("(A:12,B:12)", 0.02),
    ("(((A:3,B:3):4,C:7):6,(D:5,(E:2,F:2):3):8)", 0.015),
])
def test_k4_replay_matches_oracle_draw_for_draw(newick, lam):
    p = make(newick, lam)
    g = p.make_gpu()
    n = 200
    R = p.ranges[3] - p.ranges[2] + 1
    u = np.random.RandomState(7).random_sample(R * n * (p.tree.n_nodes - 1))
    cd = g.conditional_distribution(n, uniforms=u)
    ref = oracle.conditional_distribution(p.otree, p.oracle_mats(), p.ranges, n, uniforms=u)
    assert cd.shape == ref.shape
    assert np.all(np.diff(cd, axis=1) >= 0)
    big = ref > 1e-290
    assert rel_err(cd[big], ref[big]).max() < 1e-11
    assert np.abs(cd[~big] - ref[~big]).max() <= 1e-290 if (~big).any() else True
    g.close()


def test_k4_replay_with_lambda_mu_and_ratchet():
    # larger lambda: simulated sizes spread, the range.max ratchet (conditional_distribution.cpp:29) bites
    p = make(EXAMPLE_TREE, 0.008, mu=0.009, ranges=(0, 90, 1, 30))
    g = p.make_gpu()
    n = 150
    R = 30
    u = np.random.RandomState(11).random_sample(R * n * (p.tree.n_nodes - 1))
    cd = g.conditional_distribution(n, uniforms=u)
    ref = oracle.conditional_distribution(p.otree, p.oracle_mats(), p.ranges, n, uniforms=u)
    big = ref > 1e-290
    assert rel_err(cd[big], ref[big]).max() < 1e-11
    g.close()


def test_k4_device_rng_is_statistically_equivalent():
    from scipy import stats
    p = make(EXAMPLE_TREE, 0.004)
    g = p.make_gpu()
    n = 2000
    cd = g.conditional_distribution(n, uniforms=None, seed=12345)
    assert np.all(np.diff(cd, axis=1) >= 0)
    oracle.srand(99)
    ref = oracle.conditional_distribution(p.otree, p.oracle_mats(), p.ranges, n)
    pv = []
    for r in (0, 3, 7, 14):
        # logs rounded to 1e-8: at small root sizes a sizeable share of the draws is ONE family (every leaf empty), a block of equal
        # values in both samples; the GPU's and the oracle's value of it agree to ~1e-13, and unrounded the statistic would see the
        # whole block as a gap between the two distributions
        a = np.round(np.log(np.maximum(cd[r], 1e-300)), 8)
        b = np.round(np.log(np.maximum(ref[r], 1e-300)), 8)
        pv.append(stats.ks_2samp(a, b).pvalue)
    assert min(pv) > 1e-4, pv
    # a different seed gives a different sample
    cd2 = g.conditional_distribution(n, uniforms=None, seed=54321)
    assert not np.array_equal(cd, cd2)
    g.close()


@pytest.mark.parametrize("with_err", [False, True])
def test_k5_family_pvalues_match_oracle(with_err):
    err = None
    ranges = (0, 62, 1, 15)
    if with_err:
        dim = 63
        E = np.zeros((dim, dim))
        eps = 0.03
        for j in range(dim):
            for d, v in ((-1, eps), (0, 1 - 2 * eps), (1, eps)):
                if 0 <= j + d < dim:
                    E[j + d, j] = v
        E[0, 0] = 1 - eps
        E[dim - 1, dim - 1] = 1 - eps
        err = {k: E for k in range(5)}
    p = make(EXAMPLE_TREE, 0.004, F=40, err=err, ranges=ranges)
    p.counts[0, :] = 0  # an all-zero family: empty root range -> p-value 0 (viterbi.cpp:32-39)
    g = p.make_gpu()
    n = 400
    R = ranges[3] - ranges[2] + 1
    u = np.random.RandomState(3).random_sample(R * n * (p.tree.n_nodes - 1))
    cd = oracle.conditional_distribution(p.otree, p.oracle_mats(), p.ranges, n, uniforms=u)
    pv = g.pvalues(cd)
    ref = np.array([oracle.family_pvalue(p.otree, p.oracle_mats(), p.counts[f], cd, leaf_err=p.oracle_leaf_err())[0]
                    for f in range(len(p.counts))])
    assert pv[0] == 0.0 and ref[0] == 0.0
    assert np.abs(pv - ref).max() <= 1.0 / n + 1e-12
    assert (pv == ref).mean() > 0.97
    g.close()


def test_k4_then_score_still_correct():
    # the conditional distribution reuses the vector slots; a following score must be unaffected
    p = make(EXAMPLE_TREE, 0.004, F=33)
    g = p.make_gpu()
    s0, _ = g.score()
    g.conditional_distribution(50, uniforms=None, seed=1)
    s1, _ = g.score()
    assert s0 == s1
    o = p.oracle_score(want_L=False)
    assert abs(s1 - o["score"]) < 1e-6
    g.close()


def test_k4_rows_split_equals_whole_distribution():
    # cafe_gpu_conditional_distribution_rows: the rows two ranks would compute, concatenated, are the single-rank result
    import oracle
    from util import EXAMPLE_TREE, Problem
    rng = np.random.RandomState(5)
    counts = np.maximum(0, rng.randint(1, 20, size=(16, 1)) + rng.randint(-2, 3, size=(16, 5))).astype(np.int32)
    p = Problem(EXAMPLE_TREE, counts, 0.006)
    g = p.make_gpu()
    whole = g.conditional_distribution(64, seed=11)
    R = whole.shape[0]
    cut = R // 2 + 1
    parts = np.concatenate([g.conditional_distribution_rows(64, 0, cut, seed=11), g.conditional_distribution_rows(64, cut, R, seed=11)])
    assert np.array_equal(parts, whole)
    assert g.conditional_distribution_rows(64, 3, 3, seed=11).shape == (0, 64)
    g.close()


def test_k4_k5_config5_full_size_properties():
    # BASELINE configs[4]: the conditional distribution with 1000 draws per root size (R = 500 rows at max size 400, 50 taxa) and
    # family p-values for 200 k families.  Too big for the CPU oracle as a whole: (a) rows ascending, finite, in [0, 1];
    # (b) the rows two ranks would compute equal the whole; (c) the device generator gives the same distribution again for the
    # same seed and a different one for another seed; (d) a family sample's p-values against the oracle fed with the GPU's own
    # distribution and matrices; (e) p-values do not depend on the position of a family in the table.
    from cafe_b200 import synth
    nw = synth.random_tree(50, 1)
    counts, lam0 = synth.simulate_table(nw, 200000, 400, seed=11)
    p = Problem(nw, counts, lam0, prior_lambda=60.0, ranges=(0, 480, 1, 500))
    g = p.make_gpu()
    n = 1000
    cd = g.conditional_distribution(n, seed=7)
    assert cd.shape == (500, n)
    assert np.isfinite(cd).all() and cd.min() >= 0 and cd.max() <= 1 and (np.diff(cd, axis=1) >= 0).all()
    assert (cd[:12, -1] > 0).all()     # (with 50 leaves the probability of any one leaf pattern underflows to 0 at large root sizes)
    rows = np.concatenate([g.conditional_distribution_rows(n, 0, 250, seed=7), g.conditional_distribution_rows(n, 250, 500, seed=7)])
    assert np.array_equal(rows, cd)
    assert not np.array_equal(g.conditional_distribution_rows(n, 100, 102, seed=8), cd[100:102])
    pv = g.pvalues(cd)
    assert pv.shape == (200000,) and pv.min() >= 0 and pv.max() <= 1
    assert len(np.unique(pv)) > 10 and (pv < 0.01).mean() < 0.1   # families simulated from the model itself: not piled up at 0
    mats = [None if v == p.otree.root else g.get_matrix(v) for v in range(p.otree.n_nodes)]
    idx = np.random.RandomState(3).choice(len(counts), 16, replace=False)
    ref = np.array([oracle.family_pvalue(p.otree, mats, counts[f], cd)[0] for f in idx])
    assert np.abs(pv[idx] - ref).max() <= 1.0 / n + 1e-12
    assert (pv[idx] == ref).mean() >= 0.8
    g.close()
    perm = np.random.RandomState(5).permutation(len(counts))[:50000]
    g2 = Problem(nw, counts[perm], lam0, prior_lambda=60.0, ranges=(0, 480, 1, 500)).make_gpu()
    assert np.array_equal(g2.pvalues(cd), pv[perm])
    g2.close()


def test_k4_k5_fused_windowed_path_equals_per_node_kernels(monkeypatch):
    # K4 / K5 prune with per-family column windows.  The fused kernel (prune_fused2.cu, windowed mode) and the per-node
    # kernels (prune.cu) must agree: same simulated sizes (same device generator), likelihoods to summation order, and the
    # p-values they give identical up to ties.
    nw = oracle.random_tree(11, 4)
    rng = np.random.RandomState(9)
    base = rng.randint(1, 40, size=(700, 1))
    counts = np.maximum(0, base + rng.randint(-4, 5, size=(700, 11))).astype(np.int32)
    counts[5] = 0                      # an all-zero family: empty root range in the forced range of K5
    counts[6, 3] = 140                 # one leaf far above the others: wide window, most leaves deep in the tails
    p = Problem(nw, counts, 0.006, mu=0.005)
    g = p.make_gpu()
    n = 300
    cd_fused = g.conditional_distribution(n, seed=3)
    pv_fused = g.pvalues(cd_fused)
    monkeypatch.setenv("CAFE_GPU_NO_FUSED", "1")
    cd_node = g.conditional_distribution(n, seed=3)
    pv_node = g.pvalues(cd_fused)
    monkeypatch.delenv("CAFE_GPU_NO_FUSED")
    big = cd_node > 1e-290
    assert big.mean() > 0.5
    assert rel_err(cd_fused[big], cd_node[big]).max() < 1e-12
    assert np.abs(pv_fused - pv_node).max() <= 1.0 / n + 1e-12 and (pv_fused == pv_node).mean() > 0.99
    assert pv_fused[5] == 0.0
    # and against the oracle on a sample (its distribution argument is the GPU's own: the draws differ from glibc's)
    mats = [None if v == p.otree.root else g.get_matrix(v) for v in range(p.otree.n_nodes)]
    idx = [0, 1, 5, 6, 17, 123, 699]
    ref = np.array([oracle.family_pvalue(p.otree, mats, counts[f], cd_fused)[0] for f in idx])
    assert np.abs(pv_fused[idx] - ref).max() <= 1.0 / n + 1e-12
    g.close()


def _band_error_model(dim, eps, reach=1):
    E = np.zeros((dim, dim))
    for j in range(dim):
        for d in range(-reach, reach + 1):
            if 0 <= j + d < dim:
                E[j + d, j] = (1 - 2 * eps) if d == 0 else eps / reach
    return E


@pytest.mark.parametrize("reach", [1, 3, 55])
def test_k4_error_model_leaves_through_the_abi(reach, monkeypatch):
    # The C-ABI applies the leaves' error models to the SIMULATED sizes, as get_random_probabilities does on a tree that carries
    # them (test_oracle.py::test_ref_conditional_distribution_with_error_model_bitwise, part 2).  Band models (reach < 50 sizes
    # above the diagonal) run in the windowed fused kernel; reach 55 must take the per-node kernels and give the same numbers.
    ranges = (0, 90, 1, 20)
    E = _band_error_model(91, 0.03, reach)
    p = make(EXAMPLE_TREE, 0.006, mu=0.005, ranges=ranges, err={0: E, 2: E, 4: E})
    g = p.make_gpu()
    n = 120
    R = ranges[3] - ranges[2] + 1
    u = np.random.RandomState(17).random_sample(R * n * (p.tree.n_nodes - 1))
    cd = g.conditional_distribution(n, uniforms=u)
    ref = oracle.conditional_distribution(p.otree, p.oracle_mats(), p.ranges, n, uniforms=u, leaf_err=p.oracle_leaf_err())
    plain = oracle.conditional_distribution(p.otree, p.oracle_mats(), p.ranges, n, uniforms=u)
    assert not np.allclose(ref, plain, rtol=1e-6)
    big = ref > 1e-290
    assert rel_err(cd[big], ref[big]).max() < 1e-11
    monkeypatch.setenv("CAFE_GPU_NO_FUSED", "1")
    cd_node = g.conditional_distribution(n, uniforms=u)
    monkeypatch.delenv("CAFE_GPU_NO_FUSED")
    assert rel_err(cd[big], cd_node[big]).max() < 1e-12
    g.close()
