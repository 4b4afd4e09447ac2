"""Branch cutting (`report ... branchcutting`, cafe/branch_cutting.cpp:101-272) on the GPU against the oracle's restatement,
which tests/test_oracle.py::test_ref_branch_cutting_bitwise pins bit for bit against the reference's cut_branch +
compute_cutpvalues."""
import os

import numpy as np
import pytest

import oracle
from cafe_b200 import gpu as cgpu, host as chost

pytestmark = pytest.mark.gpu

TREE = "(((chimp:6,human:6):81,(mouse:17,rat:17):70):6,dog:93)"
SPECIES = ["chimp", "human", "mouse", "rat", "dog"]


def _table(tmp_path, counts):
    p = str(tmp_path / "fam.tab")
    with open(p, "w") as f:
        f.write("\t".join(["Desc", "Family ID"] + SPECIES) + "\n")
        for i, r in enumerate(counts):
            f.write("\t".join(["d", "F%d" % i] + [str(int(x)) for x in r]) + "\n")
    return p


def _counts():
    rs = np.random.RandomState(4)
    c = rs.poisson(6, size=(14, 5)).astype(np.int32)
    c[3] = [20, 1, 0, 2, 1]       # an outlier family (the table's maximum: ranges 0..70 / 1..30)
    c[9] = c[2]                   # a duplicate: takes its first occurrence's values
    c[11] = [0, 0, 1, 0, 0]
    return c


def _expected(t, lam_node, mu_node, counts, ranges, N, maxp, cutoff, seed):
    oracle.srand(seed)
    rows = []
    for b in range(t.n_nodes):
        if b == t.root:
            rows.append(np.full(len(counts), -1.0))   # cafe_branch_cutting, branch_cutting.cpp:236-240
            continue
        rows.append(oracle.branch_cut(t, lam_node, mu_node, counts, ranges, N, b, maxp, cutoff)["pvalues"])
    return np.array(rows)


@pytest.mark.parametrize("mu_ratio", [0.0, 0.5])
def test_branch_cutting_through_the_host_mirror_matches_the_oracle(tmp_path, mu_ratio):
    """Every branch of the 5-taxon tree, distributions replayed from the rand() stream after `seed 10` like the single-threaded
    reference: one-sided cuts (a leaf or the root's other child is cut off) and two-sided cuts, filtered families, a duplicate."""
    counts = _counts()
    path = _table(tmp_path, counts)
    N = 40
    lam = 0.006
    s = chost.Session(quiet=True)
    assert s.command("load -i %s -t 1 -r %d -p 0.05" % (path, N)) == 0
    assert s.command("tree " + TREE) == 0
    if mu_ratio > 0:
        assert s.command("lambdamu -l %.10g -m %.10g" % (lam, lam * mu_ratio)) == 0
    else:
        assert s.command("lambda -l %.10g" % lam) == 0
    rg = s.ranges()
    ranges = (rg["min"], rg["max"], rg["root_min"], rg["root_max"])
    assert ranges == (0, 70, 1, 30)
    rs = np.random.RandomState(8)
    maxp = rs.uniform(0, 0.05, len(counts))
    maxp[[1, 6]] = [0.3, 0.051]               # filtered: -1 on every branch
    s.set_max_pvalues(maxp)
    assert s.command("seed 10") == 0
    cut = s.branch_cutting(N)
    t = oracle.parse_newick(TREE)
    lam_node = np.full(t.n_nodes, lam)
    mu_node = np.full(t.n_nodes, lam * mu_ratio if mu_ratio > 0 else -1.0)
    # the reference skips duplicates (ref != i) and copies them afterwards: evaluate the unique rows, then copy
    exp = _expected(t, lam_node, mu_node, counts, ranges, N, maxp, 0.05, 10)
    exp[:, 9] = exp[:, 2]
    assert cut.shape == exp.shape == (t.n_nodes, len(counts))
    assert (cut[:, [1, 6]] == -1).all() and (cut[t.root] == -1).all()
    # p-values are ranks / sums of ranks: equal unless a likelihood sits within round-off of a simulated one
    diff = np.abs(cut - exp)
    assert diff.max() <= 1.0 / (N // 10) / (N // 10) + 1e-12, (diff.max(), np.argwhere(diff > 1e-12))
    assert (diff <= 1e-12).mean() > 0.97
    s.close()


def test_cut_pvalues_kernels_against_a_direct_evaluation():
    """cafe_gpu_cut_pvalues on synthetic rows: both kernels against the formulas of branch_cutting.cpp:20-44 / :125-147 evaluated
    with the oracle's pvalue(), including a zero in the second distribution (division by zero -> inf -> rank above every entry)."""
    rs = np.random.RandomState(1)
    F, rf, n = 7, 9, 12
    L1 = rs.uniform(0, 1, (F, rf)) * 10.0 ** rs.randint(-12, 0, (F, rf))
    L2 = rs.uniform(0, 1, (F, rf)) * 10.0 ** rs.randint(-12, 0, (F, rf))
    cd1 = np.sort(rs.uniform(0, 1, (rf, n)) * 10.0 ** rs.randint(-20, 0, (rf, n)), axis=1)
    cd2 = np.sort(rs.uniform(0, 1, (rf, n)) * 10.0 ** rs.randint(-12, 0, (rf, n)), axis=1)
    cd2[3, 0] = 0.0
    g = cgpu.CafeGpu(0)
    one = g.cut_pvalues(L1, cd1)
    two = g.cut_pvalues(L1, cd1, L2, cd2)
    g.close()
    exp1 = np.array([max(oracle.pvalue(L1[f, s], cd1[s]) for s in range(rf)) for f in range(F)])
    assert np.array_equal(one, exp1)
    exp2 = np.zeros(F)
    with np.errstate(divide="ignore", invalid="ignore"):
        for f in range(F):
            for s2 in range(rf):
                for s1 in range(rf):
                    p = 0.0
                    for t in range(n):
                        p += oracle.pvalue(float(np.float64(L1[f, s1]) * np.float64(L2[f, s2]) / np.float64(cd2[s2, t])), cd1[s1])
                    exp2[f] = max(exp2[f], p / n)
    assert np.array_equal(two, exp2)


def test_report_branchcutting_writes_the_cut_pvalue_column(tmp_path):
    """`report <name> branchcutting`: the family lines carry the cut p-values of every node in id order ('-' for -1) before the
    likelihood-ratio column, and - the reference's inverted header flags, reports.cpp:456-457 - the 'cut P-value' title
    disappears from the header exactly when the column is present."""
    counts = _counts()
    path = _table(tmp_path, counts)
    s = chost.Session(quiet=True)
    assert s.command("load -i %s -t 1 -r 40 -p 0.9" % path) == 0
    assert s.command("tree " + TREE) == 0
    assert s.command("lambda -l 0.006") == 0
    assert s.command("seed 10") == 0
    assert s.command("report %s branchcutting" % (tmp_path / "rep")) == 0
    lines = open(str(tmp_path / "rep") + ".cafe").read().split("\n")
    hdr = next(i for i, ln in enumerate(lines) if ln.startswith("'ID'"))
    assert lines[hdr] == "'ID'\t'Newick'\t'Family-wide P-value'\t'Viterbi P-values'\t'Likelihood Ratio'"
    fam = [ln for ln in lines[hdr + 1:] if ln]
    assert len(fam) == len(counts)
    maxp = s.max_pvalues()
    for i, ln in enumerate(fam):
        f = ln.split("\t")
        assert f[0] == "F%d" % i and len(f) == 6 and f[5] == ""
        vals = f[4].strip("()").split(",")
        assert len(vals) == 9 and vals[7] == "-"            # node 7 is the root
        if maxp[i] > 0.9:
            assert all(v == "-" for v in vals)
        else:
            assert all(0.0 <= float(v) <= 1.0 for k, v in enumerate(vals) if k != 7)
    assert fam[9].split("\t")[4] == fam[2].split("\t")[4]   # the duplicate family copies its first occurrence
    s.close()
