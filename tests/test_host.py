"""CPU: the C++ host mirror (cafe_b200/host) against the oracle, the compiled reference and the goldens.
Nothing here touches a GPU: these are the parts of the reference's interface that sit above the C-ABI."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle
from cafe_b200 import host as chost

DP = C.POINTER(C.c_double)
IP = C.POINTER(C.c_int)
GOLD = os.path.join(os.path.dirname(__file__), "golden")
EX_TREE = "(((chimp:6,human:6):81,(mouse:17,rat:17):70):6,dog:93)"
TREES = ["((A:1,B:1):1,(C:1,D:1):1);", EX_TREE, "(A:3,B:3)", "((A:0.9,B:3.5):2,C:5.5)",
         "((((cat:68,horse:68):4,cow:73):20,(((((chimp:4,human:4):6,orang:11):2,gibbon:13):7,(macaque:4,baboon:4):16):16,marmoset:36):57):38,(rat:36,mouse:36):96)",
         oracle.random_tree(20, 1), oracle.random_tree(50, 2)]


def test_math_bitwise_vs_oracle():
    H = chost.load_library()
    L = oracle.lib()
    rng = np.random.RandomState(0)
    for a in np.r_[rng.uniform(0.5, 900, 300), np.arange(1, 60)]:
        assert H.cafe_host_gammaln(a) == L.orc_gammaln(a)
    for n, r in rng.randint(0, 700, size=(400, 2)):
        a, b = H.cafe_host_chooseln(float(n), float(r)), L.orc_chooseln(float(n), float(r))
        assert a == b or (np.isnan(a) and np.isnan(b))
    for x in range(0, 80):
        assert H.cafe_host_poisspdf(x, 9.44) == L.orc_poisspdf(x, 9.44)
    assert np.array_equal(chost.lnc_table(60), oracle.lnc_table(60), equal_nan=True)
    for mx in (0, 1, 10, 34, 100, 200, 400):
        a, b, c, d = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        L.orc_init_family_size(mx, C.byref(a), C.byref(b), C.byref(c), C.byref(d))
        r = chost.init_family_size(mx)
        assert (r["root_min"], r["root_max"], r["min"], r["max"]) == (a.value, b.value, c.value, d.value)


def test_chi2cdf_bitwise_vs_oracle():
    # the series-only incomplete gamma of the likelihood-ratio test (libcommon/mathfunc.c:128-151), incl. the never-converging tail
    rng = np.random.RandomState(2)
    for x in np.r_[rng.uniform(0, 40, 200), 1e-300, 1e-9, 3.841458820694124, 500.0, 1999.0, 2100.0, 1e5]:
        assert chost.chi2cdf(x, 1) == oracle.chi2cdf(x, 1)
    assert abs(chost.chi2cdf(3.841458820694124, 1) - 0.95) < 1e-6


def test_pvalue_bitwise_vs_oracle():
    H = chost.load_library()
    rng = np.random.RandomState(1)
    for _ in range(300):
        n = rng.randint(1, 40)
        cd = np.sort(rng.choice(np.arange(1, 12) / 7.0, size=n))
        v = float(rng.choice(np.r_[cd, rng.uniform(0, 2, 3)]))
        assert H.cafe_host_pvalue(v, cd.ctypes.data_as(DP), n) == oracle.pvalue(v, cd)


@pytest.mark.parametrize("newick", TREES)
def test_tree_numbering_matches_reference_order(newick):
    t = chost.parse_tree(newick)
    o = oracle.parse_newick(newick)
    assert list(t.left) == list(o.left) and list(t.right) == list(o.right) and list(t.parent) == list(o.parent)
    assert t.names == o.names and t.root == o.root
    np.testing.assert_array_equal(t.branchlength, o.branchlength)
    assert all((i % 2 == 0) == (t.left[i] < 0) for i in range(t.n_nodes))  # leaves even, internal odd


def test_tree_errors():
    with pytest.raises(chost.CafeHostError, match="binary"):
        chost.parse_tree("(A:1,B:1,C:1)")
    with pytest.raises(chost.CafeHostError, match="Unbalanced"):
        chost.parse_tree("((A:1,B:1):1,C:2")


def test_lambda_tree_labels():
    m, ids = chost.parse_lambda_tree(EX_TREE, "(((2,2)1,(1,1)1)1,1)")
    assert m == 2
    # nlist order: chimp, (chimp,human), human, ((..),(..)), mouse, (mouse,rat), rat, root, dog
    assert list(ids) == [1, 0, 1, 0, 0, 0, 0, -2, 0]
    with pytest.raises(chost.CafeHostError, match="not totally specified"):
        chost.parse_lambda_tree(EX_TREE, "(((2,2),(1,1)1)1,1)")
    with pytest.raises(chost.CafeHostError, match="different topology"):
        chost.parse_lambda_tree(EX_TREE, "((1,1)1,1)")


def write_table(path, species, rows, sep="\t", ids=None):
    with open(path, "w") as f:
        f.write(sep.join(["Desc", "Family ID"] + species) + "\n")
        for i, r in enumerate(rows):
            f.write(sep.join(["d%d" % i, ids[i] if ids else "ID%d" % i] + [str(x) for x in r]) + "\n")


def test_family_loader_and_dedup(tmp_path, ref_lib):
    rng = np.random.RandomState(2)
    rows = rng.randint(0, 6, size=(300, 4))
    rows[17] = rows[3]; rows[250] = rows[3]; rows[99] = rows[98]
    p = str(tmp_path / "fam.txt")
    write_table(p, ["A", "B", "C", "D"], rows)
    counts, ref, mx = chost.load_families(p)
    assert np.array_equal(counts, rows) and mx == rows.max()
    # reference reader + its O(F^2) duplicate detection
    ns, nf, ms = C.c_int(), C.c_int(), C.c_int()
    rc = np.zeros_like(counts); rr = np.zeros(len(rows), dtype=np.int32)
    assert ref_lib.refshim_load_families(p.encode(), -1, C.byref(ns), C.byref(nf), rc.ctypes.data_as(IP), rc.size,
                                         rr.ctypes.data_as(IP), C.byref(ms)) == 0
    assert np.array_equal(rc, counts) and np.array_equal(rr, ref) and ms.value == mx
    assert ref[17] == 3 and ref[250] == 3 and ref[3] == 3
    # -max_size filter and csv
    counts2, _, _ = chost.load_families(p, max_size=3)
    assert np.array_equal(counts2, rows[rows.max(axis=1) <= 3])
    pc = str(tmp_path / "fam.csv")
    write_table(pc, ["A", "B", "C", "D"], rows[:10], sep=",")
    assert np.array_equal(chost.load_families(pc)[0], rows[:10])


def test_error_model_reader_matches_reference_golden(tmp_path):
    z = np.load(os.path.join(GOLD, "errmodel.npz"))
    p = str(tmp_path / "errormodel.txt")
    open(p, "w").write(str(z["text"]))
    E, fd, td = chost.read_errormodel(p, 140)
    assert (fd, td) == (int(z["fromdiff"]), int(z["todiff"]))
    assert np.array_equal(E, z["E"])


def test_error_model_reader_vs_reference_variants(tmp_path, ref_lib):
    # short files are extended to range.max by copying the last row down; first/last columns absorb the remainder.
    # (A gap in the middle of the file makes the reference's reader spin forever — error_model.cpp:176-185
    # increments i instead of j — so gaps are not part of the parity surface.)
    txt = "maxcnt:12\ncntdiff -1 0 1\n0 0.0 0.8 0.2\n1 0.1 0.7 0.2\n2 0.15 0.7 0.15\n3 0.2 0.6 0.2\n"
    p = str(tmp_path / "e.txt")
    open(p, "w").write(txt)
    for rmax in (12, 30):
        E, fd, td = chost.read_errormodel(p, rmax)
        f1, t1 = C.c_int(), C.c_int()
        dim = ref_lib.refshim_read_errormodel(p.encode(), rmax, None, C.byref(f1), C.byref(t1))
        Er = np.zeros((dim, dim))
        ref_lib.refshim_read_errormodel(p.encode(), rmax, Er.ctypes.data_as(DP), C.byref(f1), C.byref(t1))
        assert E.shape == Er.shape and np.array_equal(E, Er)


def rosen(x):
    return 100 * (x[1] - x[0] ** 2) ** 2 + (1 - x[0]) ** 2


def walled(x):  # +inf beyond a wall, like the lambda*t >= 1 wall of the likelihood
    if x[0] >= 0.0107527 or x[0] < 0:
        return float("inf")
    return (x[0] - 0.02) ** 2 * 1e6


def walled2(x):
    if x[0] >= 0.3 or x[1] >= 0.25 or min(x) < 0:
        return float("inf")
    return (x[0] - 0.5) ** 2 + (x[1] - 0.1) ** 2 + x[0] * x[1]


@pytest.mark.parametrize("fn,x0", [(rosen, [-1.2, 1.0]), (rosen, [0.0, 0.0]), (walled, [0.0065]), (walled, [0.0105]),
                                   (walled2, [0.2, 0.2]), (walled2, [0.29, 0.01]), (lambda x: abs(x[0] - 3) + 1, [0.0])])
def test_fminsearch_same_path_as_reference(ref_lib, fn, x0):
    calls_a, calls_b = [], []

    def fa(x):
        calls_a.append(tuple(x)); return fn(x)

    xa, fva, ita = chost.fminsearch(fa, x0)
    n = len(x0)

    def _cb(xp, _):
        xs = [xp[i] for i in range(n)]
        calls_b.append(tuple(xs)); return float(fn(xs))

    cb = oracle.MATH_FUNC(_cb)
    x0a = np.array(x0, dtype=np.float64); xo = np.zeros(n); fo = C.c_double(); it = C.c_int()
    ref_lib.refshim_fminsearch(cb, None, n, x0a.ctypes.data_as(DP), 1e-6, 1e-6, xo.ctypes.data_as(DP), C.byref(fo), C.byref(it))
    assert calls_a == calls_b          # every vertex, in order, bit for bit
    assert np.array_equal(xa, xo) and fva == fo.value and ita == it.value


def test_prior_fit_bitwise_vs_reference(tmp_path, ref_lib):
    z = np.load(os.path.join(GOLD, "example.npz"))
    species = [str(s) for s in z["species_leaf_order"]]
    p = str(tmp_path / "ex.tab")
    write_table(p, species, z["counts"])
    s = chost.Session(quiet=True)
    assert s.command("load -i %s -t 1" % p) == 0
    assert s.command("tree " + EX_TREE) == 0
    assert s.num_families() == len(z["counts"])
    rg = s.ranges()
    assert (rg["min"], rg["max"], rg["root_min"], rg["root_max"]) == tuple(int(x) for x in z["ranges"])
    chost.srand(10)
    lam, it, sc = s.find_poisson_lambda()
    assert lam == float(z["poisson_lambda"])      # same glibc rand() start, same Nelder–Mead path
    assert np.array_equal(chost.prior_poisson(rg["root_min"], lam, 1000), z["prior"])
    s.close()


def test_session_species_mapping_and_errors(tmp_path):
    p = str(tmp_path / "t.tab")
    write_table(p, ["Dog", "Chimp", "Human", "Mouse", "Rat"], [[1, 2, 3, 4, 5], [1, 2, 3, 4, 5], [0, 0, 1, 0, 0]])
    s = chost.Session(quiet=True)
    assert s.command("tree " + EX_TREE) == 0
    assert s.command("load -i %s" % p) == 0
    counts, ref, index = s.family_table(5)
    assert list(index) == [8, 0, 2, 4, 6]          # species column -> nlist node (case-insensitive names)
    assert list(ref) == [0, 0, 2]
    # the likelihood commands need the CUDA library: without a device they fail loudly, no CPU fallback
    import torch
    if not torch.cuda.is_available():
        assert s.command("lambda -l 0.002") != 0
    assert s.command("load -i /nonexistent/file") != 0
    s.close()
    s2 = chost.Session(quiet=True)
    assert s2.command("tree (A:1,B:1,C:1)") != 0     # not binary
    assert s2.command("lambda -s") != 0              # no table, no tree
    s2.close()


@pytest.mark.parametrize("golden,with_lr", [("report_plain.cafe", False), ("report_likelihood.cafe", True)])
def test_text_report_writer_with_oracle_numbers_equals_the_stock_binary(tmp_path, golden, with_lr):
    # cafe_report_text (host/reports.cpp) fed with the ORACLE's Viterbi reconstruction, branch p-values (forced per-family ranges)
    # and the golden family p-values / likelihood ratios must reproduce, character for character, the report the stock reference
    # binary wrote for example_data.tab (tests/golden/make_golden.py): pins the writer and, once more, the oracle's Viterbi.
    z = np.load(os.path.join(GOLD, "example.npz"))
    cdz = np.load(os.path.join(GOLD, "cond_dist.npz"))
    lrz = np.load(os.path.join(GOLD, "lrt.npz"))
    counts = z["counts"]
    p = str(tmp_path / "example_data.tab")
    write_table(p, [str(s) for s in z["species_leaf_order"]], counts, ids=[str(i) for i in z["ids"]])
    t = oracle.parse_newick(EX_TREE)
    n = t.n_nodes
    ranges = tuple(int(x) for x in z["ranges"])
    mats = oracle.node_matrices(t, [0.005] * n, [-1.0] * n, max(ranges[1], ranges[3]))
    sizes = np.zeros((len(counts), n), dtype=np.int32)
    bpv = np.zeros((len(counts), n))
    for f, c in enumerate(counts):
        fr = oracle.forced_range(c)
        sizes[f], _ = oracle.viterbi(t, mats, c, fr)
        bpv[f] = oracle.viterbi_branch_pvalues(t, mats, sizes[f], fr[1])
    s = chost.Session(quiet=True)
    assert s.command("load -i %s -t 1 -p 0.05 -r 100" % p) == 0
    assert s.command("tree " + EX_TREE) == 0
    out = str(tmp_path / "out.cafe")
    s.report_text_from([0.005], sizes, bpv, cdz["pvalues"], out, likelihood_ratios=lrz["ratios_stock"] if with_lr else None)
    s.close()
    assert open(out).read() == open(os.path.join(GOLD, golden)).read()
