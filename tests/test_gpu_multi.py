"""-m gpu, needs >= 2 devices (skipped otherwise): multi-GPU behind the C-ABI (include/cafe_gpu.h "Multi-GPU", csrc/comm.cu).

(a) one process, several devices: cafe_gpu_create_multi shards the families, K1 is sharded over the devices, the matrices are
    all-gathered with NCCL inside the library, the score is reduced in rank order;
(b) one process per GPU: cafe_gpu_comm_unique_id / cafe_gpu_comm_init, the id handed over through a file;
(c) the C++ host (cafe_gpu_shell, CAFE_GPUS=2) runs `lambda -s` on two GPUs with the same simplex path as on one.
The single-device context is the oracle-checked baseline (tests/test_gpu_parity.py); here the two must agree with it:
per-family results bit for bit (the kernels and their inputs are the same), the score to summation order."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def n_devices():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


needs2 = pytest.mark.skipif("n_devices() < 2", reason="needs two CUDA devices")


def problem(F=300, seed=3):
    from util import Problem, random_tree, simulate_families
    import oracle
    nw = random_tree(9, 5)
    t = oracle.parse_newick(nw)
    lam = 0.01
    roots = 1 + np.random.RandomState(seed).poisson(6.0, F)
    counts = simulate_families(t, [lam] * t.n_nodes, [-1.0] * t.n_nodes, 120, F, roots, seed)
    lam_node = np.where(np.arange(t.n_nodes) % 3 == 0, 0.012, 0.008)   # several distinct keys
    p = Problem(nw, counts, lam)
    p.lam_node = lam_node
    p.mu_node = np.where(np.arange(t.n_nodes) % 2 == 0, 0.009, 0.007)
    return p


@needs2
def test_multi_device_context_matches_single_device():
    p = problem()
    one = p.make_gpu()
    s1, z1 = one.score()
    lp1, ml1, am1 = one.family_results()
    L1 = one.family_likelihoods()
    two = p.make_gpu(devices=[0, 1])
    assert two.num_devices() == 2
    s2, z2 = two.score()
    lp2, ml2, am2 = two.family_results()
    assert z1 == z2 == -1
    assert abs(s1 - s2) <= 1e-9 * abs(s1)
    assert np.array_equal(lp1, lp2) and np.array_equal(ml1, ml2) and np.array_equal(am1, am2)
    assert np.array_equal(L1, two.family_likelihoods())
    for node in (0, 3, 6):
        assert np.array_equal(one.get_matrix(node), two.get_matrix(node))
    # a second evaluation with other rates through the one-call objective
    s1b, _ = one.objective(p.lam_node * 1.1, p.mu_node * 0.9)
    s2b, _ = two.objective(p.lam_node * 1.1, p.mu_node * 0.9)
    assert abs(s1b - s2b) <= 1e-9 * abs(s1b) and s1b != s1
    # family-wide passes fan out over the devices: p-values, Viterbi, conditional distribution rows, likelihood-ratio test
    cd1 = one.conditional_distribution(40, seed=7)
    cd2 = two.conditional_distribution(40, seed=7)
    assert np.array_equal(cd1, cd2)
    assert np.array_equal(one.pvalues(cd1), two.pvalues(cd1))
    v1, v2 = one.viterbi(), two.viterbi()
    assert np.array_equal(v1[0], v2[0]) and np.array_equal(v1[1], v2[1])
    tested = (np.arange(len(p.counts)) % 7 == 2).astype(np.uint8)
    a, b = one.likelihood_ratio_test(tested), two.likelihood_ratio_test(tested)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    one.close()
    two.close()


@needs2
def test_zero_likelihood_family_is_reported_across_devices():
    p = problem()
    p.lam_node = np.full_like(p.lam_node, 0.5)   # lambda * t > 1 on every branch: every family has likelihood 0
    p.mu_node = np.full_like(p.mu_node, -1.0)
    one, two = p.make_gpu(), p.make_gpu(devices=[0, 1])
    assert one.score() == two.score() and one.score()[1] == 0
    one.close()
    two.close()


_WORKER = r"""
import sys, os, time, json
import numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
rank, world, idfile = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
from cafe_b200 import gpu as cgpu
from test_gpu_multi import problem
if rank == 0:
    open(idfile + ".tmp", "wb").write(cgpu.comm_unique_id()); os.rename(idfile + ".tmp", idfile)
while not os.path.exists(idfile):
    time.sleep(0.05)
uid = open(idfile, "rb").read()
p = problem()
F = len(p.counts); lo, hi = rank * F // world, (rank + 1) * F // world
g = p.make_gpu(comm=(uid, rank, world), lo=lo, hi=hi)
s, z = g.score()
s2, _ = g.objective(p.lam_node * 1.1, p.mu_node * 0.9)
lp, ml, am = g.family_results()
print(json.dumps({{"rank": rank, "score": s, "zero": z, "score2": s2, "ml_sum": float(ml.sum()), "n": int(len(ml))}}))
g.close()
"""


@needs2
def test_one_process_per_gpu_communicator(tmp_path):
    import json
    p = problem()
    one = p.make_gpu()
    s1, _ = one.score()
    s1b, _ = one.objective(p.lam_node * 1.1, p.mu_node * 0.9)
    one.close()
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    idfile = str(tmp_path / "nccl_id")
    procs = [subprocess.Popen([sys.executable, str(script), str(r), "2", idfile], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(2)]
    outs = []
    for pr in procs:
        o, e = pr.communicate(timeout=300)
        assert pr.returncode == 0, e[-2000:]
        outs.append(json.loads(o.strip().splitlines()[-1]))
    assert outs[0]["score"] == outs[1]["score"] and outs[0]["score2"] == outs[1]["score2"]   # the same bits on every rank
    assert abs(outs[0]["score"] - s1) <= 1e-9 * abs(s1) and abs(outs[0]["score2"] - s1b) <= 1e-9 * abs(s1b)
    assert outs[0]["n"] + outs[1]["n"] == len(p.counts)


@needs2
def test_host_shell_lambda_search_on_two_gpus(tmp_path):
    from cafe_b200 import buildlib
    z = np.load(os.path.join(ROOT, "tests", "golden", "example.npz"))
    species = [str(s) for s in z["species_leaf_order"]]
    tab = tmp_path / "example_data.tab"
    with open(tab, "w") as f:
        f.write("\t".join(["FAMILYDESC", "FAMILY"] + species) + "\n")
        for i, r in zip(z["ids"], z["counts"]):
            f.write("\t".join(["d", str(i)] + [str(x) for x in r]) + "\n")
    script = tmp_path / "run.sh"
    script.write_text("seed 10\nload -i %s -t 1\ntree (((chimp:6,human:6):81,(mouse:17,rat:17):70):6,dog:93)\nlambda -s\n" % tab)
    outs = []
    for gpus in (None, "2"):
        env = dict(os.environ)
        env.pop("CAFE_GPUS", None)
        if gpus:
            env["CAFE_GPUS"] = gpus
        r = subprocess.run([buildlib.SHELL_BIN, str(script)], capture_output=True, text=True, env=env, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(r.stdout)
    lam = [[ln for ln in o.splitlines() if "Lambda" in ln and "Score" in ln] for o in outs]
    assert len(lam[0]) > 10 and len(lam[0]) == len(lam[1])       # the same simplex path, call for call
    last = [float(ln.split("Score:")[1].split()[0]) for ln in (lam[0][-1], lam[1][-1])]
    assert abs(last[0] - last[1]) <= 1e-9 * abs(last[0])
    assert lam[0][-1].split("&")[0] == lam[1][-1].split("&")[0]  # lambda-hat printed identically


@needs2
def test_branch_cutting_spreads_branches_over_two_gpus(tmp_path):
    """`report ... branchcutting` with the device generator (more than one thread): with CAFE_GPUS=2 the branches alternate between
    the devices on two host threads; the seeds are drawn in branch order beforehand, so the report is the same file."""
    from cafe_b200 import buildlib
    rs = np.random.RandomState(4)
    counts = rs.poisson(6, size=(40, 5))
    tab = tmp_path / "fam.tab"
    with open(tab, "w") as f:
        f.write("\t".join(["Desc", "Family ID", "chimp", "human", "mouse", "rat", "dog"]) + "\n")
        for i, r in enumerate(counts):
            f.write("\t".join(["d", "F%d" % i] + [str(int(x)) for x in r]) + "\n")
    outs = []
    for gpus in (None, "2"):
        script = tmp_path / ("run%s.sh" % (gpus or "1"))
        rep = tmp_path / ("rep%s" % (gpus or "1"))
        script.write_text("seed 10\nload -i %s -t 4 -r 60 -p 0.9\ntree (((chimp:6,human:6):81,(mouse:17,rat:17):70):6,dog:93)\n"
                          "lambda -l 0.006\nreport %s branchcutting\n" % (tab, rep))
        env = dict(os.environ)
        env.pop("CAFE_GPUS", None)
        if gpus:
            env["CAFE_GPUS"] = gpus
        r = subprocess.run([buildlib.SHELL_BIN, str(script)], capture_output=True, text=True, env=env, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(open(str(rep) + ".cafe").read())
    cut = [[ln.split("\t")[4] for ln in o.split("\n") if ln.startswith("F")] for o in outs]
    assert len(cut[0]) == 40 and cut[0] == cut[1]
