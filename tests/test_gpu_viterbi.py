"""-m gpu: Viterbi ancestral reconstruction (cafe_gpu_viterbi) against the CPU oracle, which tests/test_oracle.py pins bit for
bit against the compiled reference.  Reconstructed sizes are integers: exact equality, except where two histories tie to
within rounding (then the products must agree to 1e-12 relative); the root's max-product likelihood within 1e-12 relative."""
import numpy as np
import pytest

import oracle
from cafe_b200 import host as chost

from util import EXAMPLE_TREE, Problem, random_tree, rel_err

pytestmark = pytest.mark.gpu


def _counts(n_leaves, F, hi, seed):
    rng = np.random.RandomState(seed)
    base = rng.randint(0, hi, size=(F, 1))
    return np.maximum(0, base + rng.randint(-3, 4, size=(F, n_leaves))).astype(np.int32)


def _check(p):
    g = p.make_gpu()
    sizes, ml = g.viterbi()
    g.close()
    mats = p.oracle_mats()
    le = p.oracle_leaf_err()
    n_diff = 0
    for f in range(len(p.counts)):
        s_o, ml_o = oracle.viterbi(p.otree, mats, p.counts[f], p.ranges, leaf_err=le)
        assert rel_err(ml[f], ml_o, 1e-300) <= 1e-12
        assert np.array_equal(sizes[f, 0::2], p.counts[f])
        if not np.array_equal(sizes[f], s_o):
            n_diff += 1
    # CUDA and glibc round the matrix entries differently in the last bit: a tie between two histories may break differently
    assert n_diff <= max(1, len(p.counts) // 50), n_diff
    return sizes


def test_viterbi_example_tree():
    _check(Problem(EXAMPLE_TREE, _counts(5, 64, 25, 3), 0.005))


def test_viterbi_two_classes_lambda_mu():
    _check(Problem(EXAMPLE_TREE, _counts(5, 40, 20, 4), [0.004, 0.007], mu=[0.003, 0.005], lambda_tree="(((2,2)1,(1,1)1)1,1)"))


@pytest.mark.parametrize("n_leaves,seed", [(3, 2), (13, 4), (20, 5)])
def test_viterbi_random_trees(n_leaves, seed):
    nw = random_tree(n_leaves, seed)
    _check(Problem(nw, _counts(n_leaves, 50, 30, seed), 0.01))


def test_viterbi_root_range_wider_than_vector_and_ragged_count():
    nw = random_tree(13, 4)
    _check(Problem(nw, _counts(13, 37, 40, 9), 0.008, ranges=(0, 60, 3, 97)))


def test_viterbi_error_model():
    rg = chost.init_family_size(30)
    dim = rg["max"] + 1
    E = np.zeros((dim, dim))
    eps = 0.0274
    for j in range(dim):
        for d, v in ((-1, eps), (0, 1 - 2 * eps), (1, eps)):
            if 0 <= j + d < dim:
                E[j + d, j] = v
    E[0, 0] = 1 - eps
    E[dim - 1, dim - 1] = 1 - eps
    counts = np.minimum(_counts(5, 48, 26, 8), 30)
    p = Problem(EXAMPLE_TREE, counts, 0.004, err={k: E for k in range(5)},
                ranges=(rg["min"], rg["max"], rg["root_min"], rg["root_max"]))
    _check(p)


def _check_report(p):
    """cafe_gpu_viterbi_report (forced per-family ranges + branch p-values) against the oracle."""
    g = p.make_gpu()
    sizes, pv = g.viterbi_report()
    g.close()
    mats = p.oracle_mats()
    le = p.oracle_leaf_err()
    n_diff = 0
    for f in range(len(p.counts)):
        fr = oracle.forced_range(p.counts[f])
        s_o, _ = oracle.viterbi(p.otree, mats, p.counts[f], fr, leaf_err=le)
        if not np.array_equal(sizes[f], s_o):
            n_diff += 1
            continue
        pv_o = oracle.viterbi_branch_pvalues(p.otree, mats, s_o, fr[1])
        assert pv[f, p.otree.root] == -1 and pv_o[p.otree.root] == -1
        nz = np.arange(p.otree.n_nodes) != p.otree.root
        # an entry that equals the realised transition to the last bit on the CPU may differ by an ulp on the GPU (half vs full
        # weight of one term): compare with the weight of the realised transition as the allowance
        assert np.abs(pv[f, nz] - pv_o[nz]).max() <= 1e-9
    assert n_diff <= max(1, len(p.counts) // 50), n_diff


def test_viterbi_report_forced_ranges_and_branch_pvalues():
    counts = _counts(5, 64, 25, 13)
    counts[0] = 0            # almost empty
    counts[0, 0] = 1
    counts[1] = 0            # an all-zero family: empty forced root range (root children reconstruct to 0, root to root_min)
    mx = int(counts.max())
    rg = chost.init_family_size(mx)
    _check_report(Problem(EXAMPLE_TREE, counts, 0.005, ranges=(rg["min"], rg["max"], rg["root_min"], rg["root_max"])))
    nw = random_tree(13, 4)
    counts = np.maximum(1, _counts(13, 40, 30, 14))
    rg = chost.init_family_size(int(counts.max()))
    _check_report(Problem(nw, counts, 0.01, ranges=(rg["min"], rg["max"], rg["root_min"], rg["root_max"])))


@pytest.mark.parametrize("newick,missing,ranges", [
    (EXAMPLE_TREE, [4], (0, 72, 1, 36)),        # dog: a child of the root, root range narrower than the vector
    (EXAMPLE_TREE, [0, 1], (0, 72, 1, 36)),     # a whole cherry without data
    (None, [2, 7], (0, 60, 3, 97)),             # random 13-taxon tree, root range wider than the vector
])
def test_viterbi_species_without_data(newick, missing, ranges):
    """Leaf count -1 = a species of the tree without a column in the table (familysize < 0, cafe/viterbi.cpp:236-250): all-ones
    leaf vector over the parent's range, the leaf's size reconstructed like an ancestor's.  The oracle is pinned against the
    reference on exactly this (tests/test_oracle.py::test_ref_viterbi_missing_species_bitwise).  The likelihood paths refuse
    such a table, as the reference's pruning asserts (cafe_tree.c:207)."""
    from cafe_b200 import gpu as cgpu
    nw = newick or random_tree(13, 4)
    n_leaves = nw.count(",") + 1
    counts = _counts(n_leaves, 41, 25, 11)
    counts[:, missing] = -1
    p = Problem(nw, counts, 0.006, ranges=ranges, prior_lambda=8.0)
    g = p.make_gpu()
    sizes, ml = g.viterbi()
    with pytest.raises(cgpu.CafeGpuError):
        g.score()
    with pytest.raises(cgpu.CafeGpuError):
        g.family_likelihoods()
    g.close()
    mats = p.oracle_mats()
    obs = [k for k in range(n_leaves) if k not in missing]
    n_diff = 0
    for f in range(len(counts)):
        s_o, ml_o = oracle.viterbi(p.otree, mats, counts[f], p.ranges)
        assert rel_err(ml[f], ml_o, 1e-300) <= 1e-12
        assert np.array_equal(sizes[f, 0::2][obs], counts[f][obs]) and (sizes[f, 0::2][missing] >= 0).all()
        n_diff += not np.array_equal(sizes[f], s_o)
    assert n_diff <= 1, n_diff
