"""-m gpu: the host-facing entry points (load / tree / lambda / lambdamu / pvalue / report through the
reference-named callbacks) on the GPU path against goldens produced by the unmodified reference binary
(tests/golden/make_golden.py).  Bar from BASELINE.json: lambda-hat within 1e-6 relative on
example_data.tab; scores within max(1e-6, 1e-12*|score|)."""
import os

import numpy as np
import pytest

from cafe_b200 import host as chost

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
EX_TREE = "(((chimp:6,human:6):81,(mouse:17,rat:17):70):6,dog:93)"


@pytest.fixture()
def example_table(tmp_path):
    z = np.load(os.path.join(GOLD, "example.npz"))
    species = [str(s) for s in z["species_leaf_order"]]
    p = str(tmp_path / "example_data.tab")
    with open(p, "w") as f:
        f.write("\t".join(["FAMILYDESC", "FAMILY"] + species) + "\n")
        for i, r in zip(z["ids"], z["counts"]):
            f.write("\t".join(["d", str(i)] + [str(x) for x in r]) + "\n")
    return p, z


def session(path):
    s = chost.Session(quiet=True)
    assert s.command("seed 10") == 0
    assert s.command("load -i %s -t 1" % path) == 0
    assert s.command("tree " + EX_TREE) == 0
    return s


def test_fixed_lambda_scores_and_family_likelihoods(example_table):
    path, z = example_table
    s = session(path)
    # one `lambda` command: every such command refits the Poisson prior from a fresh rand() start
    # (SURVEY.md App. C), the golden scores all use the first fit after `seed 10`
    assert s.command("lambda -l %.10g" % z["lambdas"][0]) == 0
    np.testing.assert_array_equal(s.prior(1000), z["prior"])
    for i, lam in enumerate(z["lambdas"]):
        neg = s.objective([lam])
        assert abs(-neg - z["scores"][i]) <= max(1e-6, 1e-12 * abs(z["scores"][i]))
        L = s.family_likelihoods()
        ref = z[f"L{i}"]
        big = ref > 1e-290
        assert (np.abs(L[big] - ref[big]) / ref[big]).max() < 1e-11
    # beyond the lambda*t = 1 wall every family has likelihood 0 -> +inf objective
    assert s.objective([0.0108]) == np.inf
    assert s.objective([-0.001]) == np.inf
    s.close()


def test_lambda_search_matches_reference(example_table):
    path, z = example_table
    s = session(path)
    assert s.command("lambda -s") == 0
    lam = s.parameters()[0]
    assert abs(lam - float(z["search_lambda"])) / float(z["search_lambda"]) < 1e-6
    assert s.objective_calls() == len(z["search_trace"])   # the same simplex path, vertex for vertex
    assert abs(s.objective([lam]) - float(z["search_neg_score"])) < 1e-5
    s.close()


def test_two_class_search_and_fixed_score(example_table):
    path, z = example_table
    s = session(path)
    assert s.command("lambda -s -t (((2,2)1,(1,1)1)1,1)") == 0
    lam = s.parameters()
    ref = z["search2_lambdas"]
    assert np.abs(lam - ref).max() / ref.max() < 1e-5 and abs(lam[0] - ref[0]) / ref[0] < 1e-6
    s.close()
    s = session(path)
    assert s.command("lambda -l 0.002 0.006 -t (((2,2)1,(1,1)1)1,1)") == 0
    assert abs(-s.objective([0.002, 0.006]) - float(z["two_class_score"])) < 1e-5
    s.close()


def test_lambdamu_search_matches_reference(example_table):
    path, z = example_table
    s = session(path)
    assert s.command("lambdamu -s") == 0
    lam, mu = s.parameters()
    assert abs(lam - float(z["search_lm_lambda"])) / float(z["search_lm_lambda"]) < 1e-6
    assert abs(mu - float(z["search_lm_mu"])) / float(z["search_lm_mu"]) < 1e-6
    s.close()


def test_report_pvalues_replay_matches_reference(example_table, tmp_path):
    path, z = example_table
    g = np.load(os.path.join(GOLD, "cond_dist.npz"))
    s = chost.Session(quiet=True)
    assert s.command("load -i %s -t 1 -r %d" % (path, int(g["n_samples"]))) == 0
    assert s.command("tree " + EX_TREE) == 0
    assert s.command("seed 3") == 0
    assert s.command("lambda -l %.10g" % float(g["lam"])) == 0
    assert s.command("seed 10") == 0            # the golden CD was drawn after srand(10), single thread
    assert s.command("report %s" % (tmp_path / "rep")) == 0
    cd = s.cond_dist()
    ref = g["cd"]
    assert cd.shape == ref.shape
    big = ref > 1e-290
    assert (np.abs(cd[big] - ref[big]) / ref[big]).max() < 1e-11
    pv = s.max_pvalues()
    assert np.abs(pv - g["pvalues"]).max() <= 1.0 / int(g["n_samples"]) + 1e-12
    assert (pv == g["pvalues"]).mean() > 0.95
    assert os.path.exists(str(tmp_path / "rep") + ".pvalues")
    s.close()


def test_error_model_command_changes_score_like_the_oracle(example_table, tmp_path):
    import oracle
    path, z = example_table
    e = np.load(os.path.join(GOLD, "errmodel.npz"))
    em = str(tmp_path / "errormodel.txt")
    open(em, "w").write(str(e["text"]))
    s = session(path)
    assert s.command("errormodel -all -model %s" % em) == 0
    assert s.command("lambda -l 0.005") == 0
    neg = s.objective([0.005])
    # oracle with the reference-built dense matrix on every leaf (range.max of the example table = 84 -> dim 91)
    t = oracle.parse_newick(EX_TREE)
    E, _, _ = chost.read_errormodel(em, int(z["ranges"][1]))
    ranges = tuple(int(x) for x in z["ranges"])
    mats = oracle.node_matrices(t, [0.005] * t.n_nodes, [-1.0] * t.n_nodes, max(ranges[1], ranges[3]))
    le = [E if i % 2 == 0 else None for i in range(t.n_nodes)]
    o = oracle.score(t, mats, z["counts"], ranges, s.prior(ranges[3] - ranges[2] + 1), leaf_err=le)
    assert abs(-neg - o["score"]) <= max(1e-6, 1e-12 * abs(o["score"]))
    assert abs(o["score"] - z["scores"][3]) > 1.0   # and it really differs from the error-free score
    s.close()
