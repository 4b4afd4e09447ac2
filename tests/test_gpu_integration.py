"""-m gpu: the reference's own real-data fixtures, tests/integration/test1.sh and test4.sh, through the
host mirror (cafe_b200/host: load / tree / errormodel / lambda) on the GPU path, against what the UNMODIFIED reference binary
printed for them here (tests/golden/integration.npz, made by tests/golden/make_golden_integration.py).

test1: 15 413 families x 20 taxa, `load -max_size 20` keeps 14 787, `lambda -s`: the whole simplex path — every objective
       call's lambda to the printed digits, every score to max(1e-6, 1e-12 |score|) — and lambda-hat to 1e-6 relative.
test4: 12 653 families x 13 taxa, 2 lambda classes + errormodel.txt on every leaf: the score at fixed lambdas and the root
       likelihood vectors of every 25th family (1e-11 relative)."""
import os

import numpy as np
import pytest

from cafe_b200 import host as chost

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "integration.npz"))


def write_table(path, species, table):
    with open(path, "w") as f:
        f.write("\t".join(["Description", "ID"] + [str(s) for s in species]) + "\n")
        for k, r in enumerate(table):
            f.write("\t".join([f"TEST{k}", f"TEST{k}"] + [str(int(x)) for x in r]) + "\n")


def trace_of(session_log):
    out = []
    for ln in session_log.splitlines():
        ln = ln.lstrip(".")
        if ln.startswith("Lambda : ") and "& Score:" in ln:
            a, b = ln[len("Lambda : "):].split(" & Score: ")
            out.append((a.strip(), float(b)))
    return out


def test_test1_lambda_search_follows_the_reference_call_for_call(gold, tmp_path, capfd):
    tab = str(tmp_path / "test1_families.txt")
    write_table(tab, gold["t1_species"], gold["t1_table"])
    s = chost.Session(quiet=False)
    assert s.command("seed 10") == 0
    assert s.command("tree " + str(gold["t1_newick"])) == 0
    assert s.command("load -i %s -max_size %d" % (tab, int(gold["t1_max_size"]))) == 0
    assert s.num_families() == int(gold["t1_n_families"])
    rg = s.ranges()
    assert [rg["root_min"], rg["root_max"]] == list(gold["t1_root_range"]) and [rg["min"], rg["max"]] == list(gold["t1_family_range"])
    capfd.readouterr()
    assert s.command("lambda -s") == 0
    log = capfd.readouterr().out
    ours = trace_of(log)
    ref = gold["t1_trace"]
    # the final "Lambda : x & Score: y" line of the result block is not an objective call
    calls = ours[: len(ref)]
    assert len(ours) >= len(ref) and s.objective_calls() == len(ref)
    for (lam_txt, score), (rlam, rscore) in zip(calls, ref):
        assert lam_txt == "%.14f" % rlam                       # the same simplex vertex, to the printed digits
        assert abs(score - rscore) <= max(1e-6, 1e-12 * abs(rscore))
    lam_hat = s.parameters()[0]
    assert abs(lam_hat - float(gold["t1_lambda"])) <= 1e-6 * float(gold["t1_lambda"])
    s.close()


def test_test4_two_lambda_classes_with_error_model(gold, tmp_path, capfd):
    tab = str(tmp_path / "test4_families.txt")
    write_table(tab, gold["t4_species"], gold["t4_table"])
    em = str(tmp_path / "errormodel.txt")
    open(em, "w").write(str(gold["t4_errormodel_text"]))
    s = chost.Session(quiet=False)
    assert s.command("seed 10") == 0
    assert s.command("tree " + str(gold["t4_newick"])) == 0
    assert s.command("load -i %s" % tab) == 0
    rg = s.ranges()
    assert [rg["root_min"], rg["root_max"]] == list(gold["t4_root_range"]) and [rg["min"], rg["max"]] == list(gold["t4_family_range"])
    assert s.command("errormodel -all -model %s" % em) == 0
    capfd.readouterr()
    assert s.command("lambda -l %.10g %.10g -t %s -score" % (gold["t4_lambdas"][0], gold["t4_lambdas"][1], str(gold["t4_lambda_tree"]))) == 0
    log = capfd.readouterr().out
    tr = trace_of(log)
    ref = float(gold["t4_score"])
    assert tr and abs(tr[-1][1] - ref) <= max(1e-6, 1e-12 * abs(ref))
    L = s.family_likelihoods()
    refL = gold["t4_L"]
    mine = L[gold["t4_sample"]]
    big = refL > 1e-290
    assert big.sum() > 1000
    assert (np.abs(mine[big] - refL[big]) / refL[big]).max() < 1e-11
    s.close()
