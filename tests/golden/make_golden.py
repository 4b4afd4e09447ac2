"""Generate tests/golden/*.npz from the UNMODIFIED reference compiled into oracle/_ref/ (needs
/root/reference at build time: `make -C oracle ref`).  Run from the repo root:

    python tests/golden/make_golden.py

Fixtures (small, committed):
  matrices.npz  transition matrices of a few (t, lambda, mu, max) keys          <- compute_birthdeath_rates
  example.npz   example/example_data.tab counts, tree, ranges, Poisson prior,
                per-family root likelihoods and scores at several lambdas       <- compute_tree_likelihoods / get_posterior
                + the lambda search result of the reference binary (seed 10)    <- cafe_ref `lambda -s`
  cond_dist.npz conditional distribution (seed 10, 1 thread) + family p-values  <- cafe_conditional_distribution, viterbi_section
  errmodel.npz  tests/integration/errormodel.txt as the dense matrix the reference builds, and test4-style likelihoods
"""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
DP = C.POINTER(C.c_double)
IP = C.POINTER(C.c_int)
EX_TREE = "(((chimp:6,human:6):81,(mouse:17,rat:17):70):6,dog:93)"


def d(a):
    return a.ctypes.data_as(DP)


def i(a):
    return a.ctypes.data_as(IP)


def load_table(path):
    hdr = open(path).readline().rstrip("\n").split("\t")[2:]
    rows, ids = [], []
    for ln in open(path).read().split("\n")[1:]:
        if ln.strip():
            p = ln.split("\t")
            ids.append(p[1])
            rows.append([int(x) for x in p[2:]])
    return hdr, ids, np.array(rows, dtype=np.int32)


def main():
    R = oracle.ref()
    assert R is not None, "build the reference first: make -C oracle ref"
    # ---- matrices
    keys = [(10, .02, .01, 3), (1, .01, -1, 20), (68, .006335, -1, 140), (17, .004, .003, 100), (6, .002, .002, 84), (93, .02, -1, 30), (0, .1, -1, 10)]
    out = {"keys": np.array(keys)}
    for k, (t, lam, mu, mx) in enumerate(keys):
        M = np.zeros((mx + 1, mx + 1))
        R.refshim_bd_matrix(t, lam, mu, mx, d(M))
        out[f"m{k}"] = M
    np.savez_compressed(os.path.join(OUT, "matrices.npz"), **out)

    # ---- example data
    species, ids, table = load_table(os.path.join(REF, "example", "example_data.tab"))
    t = oracle.parse_newick(EX_TREE)
    leaf_names = [n.lower() for n in t.leaf_names]
    perm = [[s.lower() for s in species].index(n) for n in leaf_names]
    counts = np.ascontiguousarray(table[:, perm])  # leaf order
    rg = np.zeros(4, dtype=np.int32)
    R.refshim_init_family_size(int(counts.max()), i(rg))
    ranges = (int(rg[2]), int(rg[3]), int(rg[0]), int(rg[1]))  # min, max, root_min, root_max
    h = R.refshim_session_new(EX_TREE.encode(), *ranges)
    R.refshim_set_families(h, len(counts), i(counts), 1)
    R.refshim_srand(10)
    it, sc = C.c_int(), C.c_double()
    pl = R.refshim_find_poisson_lambda(h, C.byref(it), C.byref(sc))
    prior = np.zeros(1000)
    R.refshim_prior_poisson(ranges[2], pl, d(prior))
    lambdas = np.array([0.0005, 0.001, 0.0017, 0.005, 0.01, 0.0107])
    ex = {"newick": np.array(EX_TREE), "species_leaf_order": np.array(leaf_names), "ids": np.array(ids), "counts": counts,
          "ranges": np.array(ranges), "prior": prior, "poisson_lambda": np.array(pl), "lambdas": lambdas}
    scores = []
    Rr = ranges[3] - ranges[2] + 1
    n = R.refshim_n_nodes(h)
    for k, lam in enumerate(lambdas):
        R.refshim_set_rates(h, d(np.full(n, lam)), d(np.full(n, -1.0)))
        R.refshim_reset_cache(h)
        L = np.zeros((len(counts), Rr))
        for f in range(len(counts)):
            R.refshim_likelihoods(h, i(np.ascontiguousarray(counts[f])), d(L[f]))
        threw = C.c_int(0)
        scores.append(R.refshim_get_posterior(h, d(prior), C.byref(threw), None, 0))
        ex[f"L{k}"] = L
    ex["scores"] = np.array(scores)
    # two lambda classes, fixed (-2008.401646)
    # ---- the reference binary's searches (seed 10)
    def run_script(lines):
        with tempfile.TemporaryDirectory() as td:
            for fn in ("example_data.tab",):
                subprocess.run(["cp", os.path.join(REF, "example", fn), td], check=True)
            open(os.path.join(td, "s.sh"), "w").write("\n".join(lines) + "\n")
            r = subprocess.run([oracle.ref_binary(), "s.sh"], cwd=td, capture_output=True, text=True)
            return r.stdout
    o1 = run_script(["seed 10", "load -i example_data.tab -t 1", f"tree {EX_TREE}", "lambda -s"])
    m = re.findall(r"Lambda : ([0-9.]+) & Score: ([0-9.]+)\nDONE", o1)
    ex["search_lambda"] = np.array(float(m[-1][0])); ex["search_neg_score"] = np.array(float(m[-1][1]))
    ex["search_trace"] = np.array([[float(a), float(b)] for a, b in re.findall(r"Lambda : ([0-9.]+) & Score: (-?[0-9.inf]+)\n\.", o1.replace("-inf", "-inf"))] or [[0, 0]])
    o2 = run_script(["seed 10", "load -i example_data.tab -t 1", f"tree {EX_TREE}", "lambda -s -t (((2,2)1,(1,1)1)1,1)"])
    m2 = re.findall(r"Lambda : ([0-9.,]+) & Score: ([0-9.]+)\nDONE", o2)
    ex["search2_lambdas"] = np.array([float(x) for x in m2[-1][0].split(",")]); ex["search2_neg_score"] = np.array(float(m2[-1][1]))
    o3 = run_script(["seed 10", "load -i example_data.tab -t 1", f"tree {EX_TREE}", "lambdamu -s"])
    m3 = re.findall(r"Lambda : ([0-9.,]+) & Score: ([0-9.]+)Mu : ([0-9.,]+) & Score: ([0-9.]+)", o3)
    ex["search_lm_lambda"] = np.array(float(m3[-1][0])); ex["search_lm_mu"] = np.array(float(m3[-1][2])); ex["search_lm_neg_score"] = np.array(float(m3[-1][1]))
    o4 = run_script(["seed 10", "load -i example_data.tab -t 1", f"tree {EX_TREE}", "lambda -l 0.002 0.006 -t (((2,2)1,(1,1)1)1,1) -score"])
    m4 = re.findall(r"Score: (-[0-9.]+)", o4)
    ex["two_class_score"] = np.array(float(m4[-1]))
    np.savez_compressed(os.path.join(OUT, "example.npz"), **ex)

    # ---- the text report of the stock binary (`report <name> [likelihood]`, reports.cpp:650-708): 100 draws per root size,
    #      family p-value cut-off 0.05, generator re-seeded right before the report
    def run_report(opt):
        with tempfile.TemporaryDirectory() as td:
            subprocess.run(["cp", os.path.join(REF, "example", "example_data.tab"), td], check=True)
            open(os.path.join(td, "s.sh"), "w").write("\n".join([
                "seed 10", "load -i example_data.tab -t 1 -p 0.05 -r 100", f"tree {EX_TREE}", "lambda -l 0.005", "seed 10",
                f"report out {opt}".strip()]) + "\n")
            subprocess.run([oracle.ref_binary(), "s.sh"], cwd=td, capture_output=True, text=True)
            return open(os.path.join(td, "out.cafe")).read()
    open(os.path.join(OUT, "report_plain.cafe"), "w").write(run_report(""))
    open(os.path.join(OUT, "report_likelihood.cafe"), "w").write(run_report("likelihood"))

    # ---- conditional distribution + family p-values at lambda = 0.005 (1 thread, seed 10)
    R.refshim_set_rates(h, d(np.full(n, 0.005)), d(np.full(n, -1.0)))
    R.refshim_reset_cache(h)
    N = 100
    R.refshim_srand(10)
    cd = np.zeros((Rr, N))
    R.refshim_cond_dist(h, 1, N, d(cd))
    R.refshim_srand(10)
    u = np.array([R.refshim_unifrnd() for _ in range(Rr * N * (n - 1))])
    pv = np.zeros(len(counts))
    for f in range(len(counts)):
        pv[f] = R.refshim_family_pvalue(h, f, d(cd), Rr, N, None, None)
    np.savez_compressed(os.path.join(OUT, "cond_dist.npz"), cd=cd, uniforms=u, pvalues=pv, lam=np.array(0.005), n_samples=np.array(N))

    # ---- branch-stretch likelihood-ratio test (cafe_likelihood_ratio_test, cafe_main.c:398-431) at lambda = 0.005 on the
    #      families whose p-value above is <= 0.05; one thread.  The shim makes the tree copy carry mu = -1 (see ref_shim.cpp).
    cutoff = 0.05
    lr = np.zeros((n, len(counts)))
    R.refshim_likelihood_ratio_test(h, d(pv), cutoff, 1, d(lr))
    lr_stock = np.zeros((n, len(counts)))   # the unmodified behaviour: lengthened branches keyed with the tree-level mu = 0
    R.refshim_likelihood_ratio_test(h, d(pv), cutoff, 0, d(lr_stock))
    np.savez_compressed(os.path.join(OUT, "lrt.npz"), ratios=lr, ratios_stock=lr_stock, max_pvalues=pv, cutoff=np.array(cutoff),
                        lam=np.array(0.005))
    R.refshim_session_free(h)

    # ---- error model file -> dense matrix (reader + column-sum fix), range.max = 140 as in test4
    em_path = os.path.join(REF, "tests", "integration", "errormodel.txt")
    fd, td_ = C.c_int(), C.c_int()
    dim = R.refshim_read_errormodel(em_path.encode(), 140, None, C.byref(fd), C.byref(td_))
    E = np.zeros((dim, dim))
    R.refshim_read_errormodel(em_path.encode(), 140, d(E), C.byref(fd), C.byref(td_))
    np.savez_compressed(os.path.join(OUT, "errmodel.npz"), E=E, fromdiff=np.array(fd.value), todiff=np.array(td_.value),
                        text=np.array(open(em_path).read()))
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
