"""-m gpu: the CUDA path (through the C-ABI) against the CPU oracle on the same seeded inputs.

Tolerances (fp64 path): transition-matrix entries 1e-12 relative (entries > 1e-300);
per-family root likelihoods 1e-11 relative; per-family log max-posterior 1e-9 absolute; total score
max(1e-6, 1e-12*|score|) absolute.  The GEMM sums in a different order than the reference's serial
j-loop, so results are not bit-identical by construction.
"""
import numpy as np
import pytest

import oracle
from cafe_b200 import host as chost

from util import EXAMPLE_TREE, Problem, random_tree, rel_err, simulate_families

pytestmark = pytest.mark.gpu

TOL_MATRIX = 1e-12
TOL_L = 1e-11
TOL_LOGPOST = 1e-9


def score_tol(s):
    return max(1e-6, 1e-12 * abs(s))


def small_counts(n_leaves, F, hi, seed):
    rng = np.random.RandomState(seed)
    base = rng.randint(1, hi, size=(F, 1))
    return np.maximum(0, base + rng.randint(-2, 3, size=(F, n_leaves))).astype(np.int32)


@pytest.mark.parametrize("t,lam,mu,maxfs", [
    (10, 0.02, 0.01, 3), (1, 0.01, -1, 20), (68, 0.006335, -1, 140), (93, 0.005, -1, 250),
    (17, 0.004, 0.003, 250), (6, 0.002, 0.002, 84), (93, 0.02, -1, 30), (0.9, 0.1, -1, 10),
    (40, 0.001, 0.0015, 500),
    (7, 0.005, 0.0, 40),   # mu = 0: log(alpha) = -inf, the NaN sums clamp to 1 (the stock likelihood-ratio test keys these)
])
def test_k1_matrix_vs_oracle(t, lam, mu, maxfs):
    # a 2-leaf tree whose two branches carry the key under test
    p = Problem(f"(A:{t},B:{t})", [[1, 1]], lam, mu=None if mu < 0 else mu, ranges=(0, maxfs, 1, maxfs))
    g = p.make_gpu()
    M = g.get_matrix(0)
    ref = oracle.bd_matrix(int(t), lam, mu, maxfs)
    assert M.shape == ref.shape
    big = ref > 1e-300
    assert rel_err(M[big], ref[big]).max() <= TOL_MATRIX
    if (~big).any():
        assert np.abs(M[~big] - ref[~big]).max() <= 1e-300
    g.close()


def test_k1_reference_kat_row_sums():
    # Appendix E of SURVEY.md: t=68, lambda=0.006335 -> M[1][0], M[1][1], M[5][5]
    p = Problem("(A:68,B:68)", [[1, 1]], 0.006335, ranges=(0, 140, 1, 140))
    g = p.make_gpu()
    M = g.get_matrix(0)
    assert abs(M[1, 0] - 0.30108052950139091) < 1e-13
    assert abs(M[1, 1] - 0.48848842624205624) < 1e-13
    assert abs(M[5, 5] - 0.19579137469510913) < 1e-13
    g.close()


def check_problem(p, expect_zero=False):
    g = p.make_gpu()
    o = p.oracle_score(want_L=True)
    s, fz = g.score()
    L = g.family_likelihoods()
    lp, ml, am = g.family_results()
    if expect_zero:
        assert fz == o["first_zero"] and fz >= 0
        assert s == -np.inf
    else:
        assert fz == -1 and o["first_zero"] == -1
        assert abs(s - o["score"]) <= score_tol(o["score"]), (s, o["score"])
        assert np.abs(lp - o["logpost"]).max() <= TOL_LOGPOST
    big = o["L"] > 1e-290
    assert rel_err(L[big], o["L"][big]).max() <= TOL_L
    if (~big).any():
        assert np.abs(L[~big] - o["L"][~big]).max() <= 1e-290
    assert rel_err(ml, o["maxlik"], 1e-290).max() <= TOL_L
    # argmax may legitimately differ only where two root sizes tie to within rounding
    diff = am != o["argmax"]
    if diff.any():
        rows = np.where(diff)[0]
        assert rel_err(o["L"][rows, am[rows]], o["L"][rows, o["argmax"][rows]]).max() < 1e-10
    g.close()
    return s


def test_k2_reference_kat_small_tree():
    # tests/test.cpp:441-474 of the reference: ((A:1,B:1):1,(C:1,D:1):1), lambda=.01, leaves 5,3,2,4
    p = Problem("((A:1,B:1):1,(C:1,D:1):1);", [[5, 3, 2, 4]], 0.01, ranges=(0, 7, 0, 7), prior_lambda=3.0)
    g = p.make_gpu()
    L = g.family_likelihoods()[0]
    kat = [0, 1.4213810941710317e-13, 2.8750146893173634e-09, 4.1190257854799189e-07, 6.7380816658820656e-07,
           2.0604688982933231e-08, 3.5778191226909485e-11, 2.9103694791175412e-14]
    assert L[0] == 0
    assert rel_err(L[1:], kat[1:]).max() < 1e-12
    g.close()


def test_k2_example_tree_single_lambda():
    p = Problem(EXAMPLE_TREE, small_counts(5, 59, 30, 3), 0.005)
    check_problem(p)


def test_k2_two_lambda_classes():
    p = Problem(EXAMPLE_TREE, small_counts(5, 40, 25, 4), [0.002, 0.006], lambda_tree="(((2,2)1,(1,1)1)1,1)")
    check_problem(p)


def test_k2_lambda_mu():
    p = Problem(EXAMPLE_TREE, small_counts(5, 40, 25, 5), 0.004, mu=0.006)
    check_problem(p)


def test_k2_zero_likelihood_family_reported():
    # lambda*t >= 1 on the dog branch zeroes its matrix (fact 9): every family scores 0
    p = Problem(EXAMPLE_TREE, small_counts(5, 8, 20, 6), 0.011, first=np.arange(8, dtype=np.int32) + 100)
    g = p.make_gpu()
    s, fz = g.score()
    assert s == -np.inf and fz == 100
    g.close()


def test_k2_fractional_branch_is_identity():
    # branch length 0.9 truncates to t=0 -> identity matrix (fact 2)
    p = Problem("((A:0.9,B:3):2,C:5)", small_counts(3, 16, 12, 7), 0.01)
    check_problem(p)


@pytest.mark.parametrize("n_leaves,seed", [(2, 1), (3, 2), (8, 3), (13, 4), (20, 5)])
def test_k2_random_trees(n_leaves, seed):
    nw = random_tree(n_leaves, seed)
    ot = oracle.parse_newick(nw)
    depth = 0
    v = 0
    while ot.parent[v] >= 0:
        depth += ot.branchlength[v]
        v = ot.parent[v]
    lam = 0.25 / depth
    n = ot.n_nodes
    counts = simulate_families(ot, [lam] * n, [-1] * n, 90, 96, np.arange(1, 40), seed)
    p = Problem(nw, counts, lam)
    check_problem(p)


def test_k2_error_model_band():
    # errormatrix[observed][true], band -1..1 (tests/integration/errormodel.txt shape)
    rg = chost.init_family_size(30)
    dim = rg["max"] + 1
    E = np.zeros((dim, dim))
    eps = 0.02744140625
    for j in range(dim):
        for d, v in ((-1, eps), (0, 1 - 2 * eps), (1, eps)):
            if 0 <= j + d < dim:
                E[j + d, j] = v
    E[0, 0] = 1 - eps
    E[dim - 1, dim - 1] = 1 - eps
    counts = small_counts(5, 48, 28, 8)
    counts = np.minimum(counts, 30)
    p = Problem(EXAMPLE_TREE, counts, 0.004, err={k: E for k in range(5)},
                ranges=(rg["min"], rg["max"], rg["root_min"], rg["root_max"]))
    check_problem(p)


def test_k2_multiplicity_and_first_index():
    counts = small_counts(5, 20, 20, 9)
    mult = np.arange(1, 21, dtype=np.int32)
    p1 = Problem(EXAMPLE_TREE, counts, 0.006, mult=mult)
    g = p1.make_gpu()
    s, _ = g.score()
    lp, _, _ = g.family_results()
    assert abs(s - float((lp * mult).sum())) <= score_tol(s)
    g.close()


def test_k2_ragged_family_count_not_multiple_of_tile():
    for F in (1, 7, 129, 257):
        p = Problem(EXAMPLE_TREE, small_counts(5, F, 22, 10 + F), 0.005)
        check_problem(p)


def test_k2_large_ranges_config2_shape_sample():
    # config-2 geometry (max size 200 -> W=251, R=250, S=251) on a small family sample
    nw = random_tree(20, 1)
    ot = oracle.parse_newick(nw)
    depth = 0
    v = 0
    while ot.parent[v] >= 0:
        depth += ot.branchlength[v]
        v = ot.parent[v]
    lam = 0.25 / depth
    n = ot.n_nodes
    counts = simulate_families(ot, [lam] * n, [-1] * n, 250, 40, np.r_[np.arange(1, 60), 150, 180, 199], 11)
    counts = np.minimum(counts, 200)
    counts[0, 0] = 200
    p = Problem(nw, counts, lam)
    assert p.ranges == (0, 250, 1, 250)
    check_problem(p)


def _tree_depth(ot):
    depth, v = 0, 0
    while ot.parent[v] >= 0:
        depth += ot.branchlength[v]
        v = ot.parent[v]
    return depth


def test_k2_root_range_wider_than_vector_lambdamu():
    # S > W with W % 4 != 0 (the geometry init_family_size produces for max size > 200), cheap sizes: W=61, R=80, S=81
    nw = random_tree(13, 4)
    ot = oracle.parse_newick(nw)
    lam = 0.25 / _tree_depth(ot)
    mu = 0.8 * lam
    counts = small_counts(13, 100, 55, 21)
    p = Problem(nw, counts, lam, mu=mu, ranges=(0, 60, 1, 80))
    check_problem(p)
    p = Problem(nw, counts, lam, mu=mu, ranges=(0, 62, 3, 97))
    check_problem(p)


def test_k2_config3_shape_lambdamu_sample():
    # config-3 geometry: max size 400 -> W=481, R=500, S=501, lambda and mu free.  Two distinct branch lengths only, so
    # that the CPU oracle's S=501 matrices (3-4 s each) stay affordable; the 50-taxon topology is covered at small S.
    nw = "(((A:3,B:3):3,(C:3,D:3):3):3,((E:3,F:3):3,G:6):3)"
    lam, mu = 0.012, 0.009
    rng = np.random.RandomState(12)
    base = np.r_[rng.randint(1, 60, 36), 300, 380, 399, 400].reshape(-1, 1)
    counts = np.clip(base + rng.randint(-3, 4, size=(40, 7)), 0, 400).astype(np.int32)
    counts[-1, 0] = 400
    p = Problem(nw, counts, lam, mu=mu)
    assert p.ranges == (0, 480, 1, 500)
    check_problem(p)


def test_k2_100_taxa_four_classes_error_model():
    # config-4 topology: 100 taxa, 4 lambda classes on branch subsets, error model on every leaf (small sizes: S=91)
    nw = random_tree(100, 1)
    ot = oracle.parse_newick(nw)
    lam0 = 0.25 / _tree_depth(ot)
    n = ot.n_nodes
    counts = simulate_families(ot, [lam0] * n, [-1] * n, 90, 40, np.arange(1, 30), 13)
    counts = np.minimum(counts, 40)
    counts[0, 0] = 40
    rg = chost.init_family_size(40)
    E = _band_error_matrix(rg["max"] + 1, 0.0274)
    p = Problem(nw, counts, lam0, err={k: E for k in range(100)})
    # what cafe_shell_set_lambdas leaves on the nodes for `lambda -t` with classes 1..4
    p.lam_node = np.array([lam0, 1.3 * lam0, 0.7 * lam0, 1.1 * lam0])[np.arange(n) % 4]
    assert p.ranges == (0, 90, 1, 50)
    check_problem(p)


def test_k2_config4_shape_error_model_sample():
    # config-4 sizes (max size 200 -> S=251) with the error model, on the 5-taxon example tree with 2 classes
    rng = np.random.RandomState(14)
    base = np.r_[rng.randint(1, 50, 28), 120, 180, 199, 200].reshape(-1, 1)
    counts = np.clip(base + rng.randint(-2, 3, size=(32, 5)), 0, 200).astype(np.int32)
    counts[-1, 0] = 200
    rg = chost.init_family_size(200)
    E = _band_error_matrix(rg["max"] + 1, 0.0274)
    p = Problem(EXAMPLE_TREE, counts, [0.002, 0.003], lambda_tree="(((2,2)1,(1,1)1)1,1)", err={k: E for k in range(5)})
    assert p.ranges == (0, 250, 1, 250)
    check_problem(p)


def _band_error_matrix(dim, eps):
    E = np.zeros((dim, dim))
    for j in range(dim):
        for d, v in ((-1, eps), (0, 1 - 2 * eps), (1, eps)):
            if 0 <= j + d < dim:
                E[j + d, j] = v
    E[0, 0] = 1 - eps
    E[dim - 1, dim - 1] = 1 - eps
    return E


# ---------------------------------------------------------------------------------------------------------
# full-size properties (BASELINE configs[1]: 50 k families x 20 taxa, max size 200): too big for the CPU oracle
# as a whole, so the checks are (a) a random family sample against the oracle, (b) the three CUDA code paths
# against each other, (c) invariance under a permutation of the families, (d) score == sum of the family terms.
# ---------------------------------------------------------------------------------------------------------
def _config2_problem(F=50000):
    from cafe_b200 import synth
    nw = synth.random_tree(20, 1)
    counts, lam0 = synth.simulate_table(nw, F, 200, seed=10)
    return nw, counts, lam0


def _family_terms(nw, counts, lam, env=None):
    import os
    saved = {}
    for k, v in (env or {}).items():
        saved[k] = os.environ.get(k)
        os.environ[k] = v
    try:
        p = Problem(nw, counts, lam, prior_lambda=8.0)
        g = p.make_gpu()
        s, fz = g.score()
        lp, ml, am = g.family_results()
        g.close()
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    return p, s, fz, lp, ml, am


def test_k2_config2_full_size_properties():
    nw, counts, lam0 = _config2_problem()
    assert counts.shape == (50000, 20) and counts.max() == 200
    p, s, fz, lp, ml, am = _family_terms(nw, counts, lam0)
    assert p.ranges == (0, 250, 1, 250) and fz == -1 and np.isfinite(lp).all()
    # (d) the score is the sum of the per-family terms (lambda.cpp:721)
    assert abs(s - float(lp.sum())) <= score_tol(s)
    # (b) first-generation fused kernel and per-node kernels: same DMMA order over K, same products -> bit-identical terms
    _, s1, _, lp1, ml1, am1 = _family_terms(nw, counts, lam0, env={"CAFE_GPU_FUSED_V1": "1"})
    assert np.array_equal(ml, ml1) and np.array_equal(am, am1)
    assert np.abs(lp - lp1).max() <= 1e-12
    _, s2, _, lp2, ml2, am2 = _family_terms(nw, counts[:4096], lam0, env={"CAFE_GPU_NO_FUSED": "1"})
    assert np.array_equal(ml[:4096], ml2) and np.array_equal(am[:4096], am2)
    assert np.abs(lp[:4096] - lp2).max() <= 1e-12
    # (c) a family's result does not depend on its position in the table (tile, group, row)
    perm = np.random.RandomState(5).permutation(len(counts))
    _, s3, _, lp3, ml3, am3 = _family_terms(nw, counts[perm], lam0)
    assert np.array_equal(ml[perm], ml3) and np.array_equal(am[perm], am3) and np.array_equal(lp[perm], lp3)
    # (a) 160 random families against the CPU oracle
    idx = np.random.RandomState(6).choice(len(counts), 160, replace=False)
    po = Problem(nw, counts[idx], lam0, prior_lambda=8.0, ranges=p.ranges)
    o = po.oracle_score(want_L=False)
    assert np.abs(lp[idx] - o["logpost"]).max() <= TOL_LOGPOST
    assert rel_err(ml[idx], o["maxlik"], 1e-290).max() <= TOL_L


@pytest.mark.gpu
def test_k2_cold_contexts_are_bit_stable():
    """A fresh context (cold scratch, TLB and L2) puts the producer warp right behind the leaf-pair gatherers at the first tile
    pair - the place where a shared completion counter once let one gatherer's lead hide the other's lag (stale rows at the
    end of a gatherer's half, rarely, and only on the first launch of a context).  Twelve cold launches, every family bit for bit."""
    nw, counts, lam0 = _config2_problem()
    _, _, _, lp0, ml0, am0 = _family_terms(nw, counts, lam0, env={"CAFE_GPU_FUSED_V1": "1"})
    for _ in range(12):
        _, _, fz, lp, ml, am = _family_terms(nw, counts, lam0)
        assert fz == -1 and np.array_equal(ml, ml0) and np.array_equal(am, am0)
        assert np.abs(lp - lp0).max() <= 1e-12


def _terms_of(make_problem, counts, env=None, want_mats=False):
    """(problem, score, first zero, log-posterior terms, max likelihoods, argmax[, GPU-built matrix per node]) of one evaluation,
    optionally under A/B environment switches."""
    import os
    saved = {}
    for k, v in (env or {}).items():
        saved[k] = os.environ.get(k)
        os.environ[k] = v
    try:
        p = make_problem(counts)
        g = p.make_gpu()
        s, fz = g.score()
        lp, ml, am = g.family_results()
        mats = None
        if want_mats:
            mats = [None if v == p.otree.root else g.get_matrix(v) for v in range(p.otree.n_nodes)]
        g.close()
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    return p, s, fz, lp, ml, am, mats


def _full_size_properties(make_problem, counts, n_v1, n_pernode, n_oracle):
    """The full-size checks of test_k2_config2_full_size_properties for any workload.  The oracle sample prunes with the
    matrices the GPU built (K1 has its own S = 501 oracle test above), so the CPU never builds S = 501 matrices here."""
    p, s, fz, lp, ml, am, mats = _terms_of(make_problem, counts, want_mats=True)
    assert fz == -1 and np.isfinite(lp).all()
    assert abs(s - float(lp.sum())) <= score_tol(s)                                   # (d) score == sum of the family terms
    _, _, _, lp1, ml1, am1, _ = _terms_of(make_problem, counts[:n_v1], env={"CAFE_GPU_FUSED_V1": "1"})
    assert np.array_equal(ml[:n_v1], ml1) and np.array_equal(am[:n_v1], am1)          # (b) the three CUDA paths agree
    assert np.abs(lp[:n_v1] - lp1).max() <= 1e-12
    _, _, _, lp2, ml2, am2, _ = _terms_of(make_problem, counts[:n_pernode], env={"CAFE_GPU_NO_FUSED": "1"})
    assert np.array_equal(ml[:n_pernode], ml2) and np.array_equal(am[:n_pernode], am2)
    assert np.abs(lp[:n_pernode] - lp2).max() <= 1e-12
    perm = np.random.RandomState(5).permutation(len(counts))                           # (c) position independence
    _, _, _, lp3, ml3, am3, _ = _terms_of(make_problem, counts[perm])
    assert np.array_equal(ml[perm], ml3) and np.array_equal(am[perm], am3) and np.array_equal(lp[perm], lp3)
    idx = np.random.RandomState(6).choice(len(counts), n_oracle, replace=False)       # (a) oracle sample
    o = oracle.score(p.otree, mats, counts[idx], p.ranges, p.prior, leaf_err=p.oracle_leaf_err())
    assert np.abs(lp[idx] - o["logpost"]).max() <= TOL_LOGPOST
    assert rel_err(ml[idx], o["maxlik"], 1e-290).max() <= TOL_L
    return p


def test_k2_config3_full_size_properties():
    # BASELINE configs[2]: 200 k families x 50 taxa, max size 400 (W=481, R=500, S=501), lambda and mu free
    from cafe_b200 import synth
    nw = synth.random_tree(50, 1)
    counts, lam0 = synth.simulate_table(nw, 200000, 400, seed=11)
    assert counts.shape == (200000, 50) and counts.max() == 400
    # prior: Poisson(60) - with the Poisson(8) of the config-2 test the prior underflows to 0 at root sizes near 400 and the family
    # that carries the table's maximum gets log(0), as it would in the reference
    p = _full_size_properties(lambda c: Problem(nw, c, lam0, mu=0.8 * lam0, prior_lambda=60.0, ranges=(0, 480, 1, 500)),
                              counts, n_v1=20000, n_pernode=2048, n_oracle=24)
    assert p.otree.n_nodes == 99


def test_k2_config4_full_size_properties():
    # BASELINE configs[3]: 100 k families x 100 taxa, max size 200, 4 lambda classes on branch subsets, error model on every leaf
    from cafe_b200 import synth
    nw = synth.random_tree(100, 1)
    counts, lam0 = synth.simulate_table(nw, 100000, 200, seed=12)
    assert counts.shape == (100000, 100) and counts.max() == 200
    E = _band_error_matrix(251, 0.0274)

    def make(c):
        q = Problem(nw, c, lam0, prior_lambda=8.0, ranges=(0, 250, 1, 250), err={k: E for k in range(100)})
        n = q.otree.n_nodes
        q.lam_node = np.array([lam0, 1.3 * lam0, 0.7 * lam0, 1.1 * lam0])[np.arange(n) % 4]   # `lambda -t` with classes 1..4
        return q

    _full_size_properties(make, counts, n_v1=20000, n_pernode=2048, n_oracle=24)


def test_k1_key_shard_two_contexts_match_unsharded():
    # cafe_gpu_set_key_shard with a collective of the caller's: two contexts stand in for two ranks, the "all-gather" of d_M
    # is two device copies through torch; cafe_gpu_matrices_exchanged transposes the received matrices into d_MT locally
    import torch
    from cafe_b200 import sharding
    counts = small_counts(5, 64, 22, 31)
    p = Problem(EXAMPLE_TREE, counts, [0.002, 0.006], lambda_tree="(((2,2)1,(1,1)1)1,1)")
    ref = p.make_gpu()
    s_ref, _ = ref.score()
    D = ref.num_keys()
    assert D >= 3
    ctxs = []
    for rank in range(2):
        g = p.make_gpu()            # builds everything once (unsharded) ...
        g.set_key_shard(rank, 2)    # ... then only its own chunk
        g.set_rates(p.lam_node, p.mu_node)
        g.build_matrices()
        with pytest.raises(Exception):
            g.score()               # not before the exchange
        ctxs.append(g)
    views = []
    for g in ctxs:
        pm, pt, dpk, kpr = g.matrix_storage()
        assert kpr == (D + 1) // 2
        g.synchronize()
        views.append(torch.as_tensor(sharding._DeviceBuffer(pm, dpk * kpr * 2), device="cuda"))
    torch.cuda.synchronize()
    chunk = views[0].numel() // 2
    views[1][:chunk].copy_(views[0][:chunk])      # rank 0's chunk -> rank 1
    views[0][chunk:].copy_(views[1][chunk:])      # rank 1's chunk -> rank 0
    torch.cuda.synchronize()
    for g in ctxs:
        g.matrices_exchanged()
        s, fz = g.score()
        assert fz == -1 and s == s_ref
        for node in (0, 2, 4):
            assert np.array_equal(g.get_matrix(node), ref.get_matrix(node))
        g.close()
    ref.close()


def test_k2_error_model_same_bits_on_both_fused_kernels():
    # error-model leaves: prune_fused2.cu gathers rows of a per-leaf matrix MTE = E x MT built once per evaluation,
    # prune_fused.cu combines the rows in its epilogue - same terms in the same order, so the results are identical
    import os
    rg = chost.init_family_size(60)
    E = _band_error_matrix(rg["max"] + 1, 0.0274)
    nw = random_tree(13, 4)
    counts = np.minimum(small_counts(13, 200, 55, 41), 60)
    res = []
    for env in (None, "CAFE_GPU_FUSED_V1"):
        if env:
            os.environ[env] = "1"
        try:
            p = Problem(nw, counts, 0.004, err={k: E for k in range(0, 13, 2)},   # every second leaf has the error model
                        ranges=(rg["min"], rg["max"], rg["root_min"], rg["root_max"]))
            g = p.make_gpu()
            s, fz = g.score()
            res.append((s, fz, g.family_likelihoods(), g.family_results()))
            g.close()
        finally:
            if env:
                os.environ.pop(env, None)
    (s2, fz2, L2, r2), (s1, fz1, L1, r1) = res
    assert fz1 == fz2 == -1
    assert np.array_equal(L1, L2) and np.array_equal(r1[1], r2[1]) and np.array_equal(r1[2], r2[2])
    assert np.abs(r1[0] - r2[0]).max() <= 1e-12 and abs(s1 - s2) <= 1e-9
    o = p.oracle_score(want_L=True)
    assert rel_err(L2[o["L"] > 1e-290], o["L"][o["L"] > 1e-290]).max() <= TOL_L
