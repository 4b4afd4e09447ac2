"""Generate tests/golden/integration.npz from the reference's own real-data fixtures, tests/integration/test1.sh and
test4.sh (SURVEY.md §7 step 4 names them as the exit test), with the UNMODIFIED reference compiled into oracle/_ref/
(`make -C oracle ref`; needs /root/reference).  Run from the repo root:

    python tests/golden/make_golden_integration.py

test1: 15 413 families x 20 taxa, `load -max_size 20` keeps 14 787, single-lambda search (`lambda -s`): the stock binary's
       whole simplex path (lambda, score per objective call), its Poisson prior fit and lambda-hat — the same numbers as the
       reference's expected transcript tests/integration/test1.t:16-69.
test4: 12 653 families x 13 taxa, two lambda classes on branch subsets + errormodel.txt on every leaf, score at fixed
       lambdas (`lambda -l 0.01 0.005 -t ... -score`), and the root likelihood vectors of every 25th family.
The family tables travel inside the .npz (int16, compressed): /root/reference does not exist on the GPU box.
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

REF_IT = "/root/reference/tests/integration"
OUT = os.path.dirname(os.path.abspath(__file__))
DP = C.POINTER(C.c_double)
IP = C.POINTER(C.c_int)


def d(a):
    return a.ctypes.data_as(DP)


def i(a):
    return a.ctypes.data_as(IP)


def load_table(path):
    lines = open(path).read().split("\n")
    hdr = lines[0].rstrip("\n").split("\t")[2:]
    rows = [[int(x) for x in ln.split("\t")[2:]] for ln in lines[1:] if ln.strip()]
    return hdr, np.array(rows, dtype=np.int32)


def script_line(path, key):
    return next(ln.strip() for ln in open(path) if ln.strip().startswith(key))


def run_ref(script_name, files):
    with tempfile.TemporaryDirectory() as td:
        for fn in files:
            subprocess.run(["cp", os.path.join(REF_IT, fn), td], check=True)
        r = subprocess.run([oracle.ref_binary(), script_name], cwd=td, capture_output=True, text=True)
        return r.stdout


def leaf_order_counts(newick, species, table):
    t = oracle.parse_newick(newick)
    perm = [species.index(n) for n in t.leaf_names]
    return t, np.ascontiguousarray(table[:, perm])


def main():
    R = oracle.ref()
    assert R is not None, "build the reference first: make -C oracle ref"
    out = {}

    # ------------------------------------------------------------------ test1: lambda -s on 14 787 families
    sh1 = os.path.join(REF_IT, "test1.sh")
    nw1 = script_line(sh1, "tree ")[5:]
    species, table = load_table(os.path.join(REF_IT, "test1_families.txt"))
    o1 = run_ref("test1.sh", ["test1.sh", "test1_families.txt"])
    trace = np.array([[float(a), float(b)] for a, b in re.findall(r"Lambda : ([0-9.]+) & Score: (-?[0-9.]+|-inf)\n\.", o1)])
    fin = re.findall(r"Lambda : ([0-9.]+) & Score: ([0-9.]+)\nDONE", o1)[-1]
    pois = re.findall(r"Poisson lambda: ([0-9.]+) & Score: ([0-9.]+)", o1)[-1]
    nfam = int(re.findall(r"The number of families is (\d+)", o1)[-1])
    rootr = [int(x) for x in re.findall(r"Root Family size : (\d+) ~ (\d+)", o1)[-1]]
    famr = [int(x) for x in re.findall(r"Family size : (\d+) ~ (\d+)", o1)[-1]]
    out.update(t1_newick=np.array(nw1), t1_species=np.array(species), t1_table=table.astype(np.int16), t1_max_size=np.array(20),
               t1_n_families=np.array(nfam), t1_root_range=np.array(rootr), t1_family_range=np.array(famr),
               t1_trace=trace, t1_lambda=np.array(float(fin[0])), t1_neg_score=np.array(float(fin[1])),
               t1_poisson_lambda=np.array(float(pois[0])), t1_poisson_score=np.array(float(pois[1])))
    print("test1:", nfam, "families, lambda-hat", fin, len(trace), "objective calls")

    # ------------------------------------------------------------------ test4: two lambda classes + error model, fixed lambdas
    sh4 = os.path.join(REF_IT, "test4.sh")
    nw4 = script_line(sh4, "tree ")[5:]
    lam_line = script_line(sh4, "lambda ")
    lam_tree = re.findall(r"-t (\S+)", lam_line)[0]
    lams = [float(x) for x in re.findall(r"-l ([0-9.]+) ([0-9.]+)", lam_line)[0]]
    species4, table4 = load_table(os.path.join(REF_IT, "test4_families.txt"))
    o4 = run_ref("test4.sh", ["test4.sh", "test4_families.txt", "errormodel.txt"])
    sc4 = float(re.findall(r"Lambda : [0-9.,]+ & Score: (-[0-9.]+)", o4)[-1])
    pois4 = re.findall(r"Poisson lambda: ([0-9.]+) & Score: ([0-9.]+)", o4)[-1]
    rootr4 = [int(x) for x in re.findall(r"Root Family size : (\d+) ~ (\d+)", o4)[-1]]
    famr4 = [int(x) for x in re.findall(r"Family size : (\d+) ~ (\d+)", o4)[-1]]
    # per-family root likelihoods of a sample through the shim (same library, same session state as the script builds)
    t4, counts4 = leaf_order_counts(nw4, species4, table4)
    ranges = (famr4[0], famr4[1], rootr4[0], rootr4[1])
    h = R.refshim_session_new(nw4.encode(), *ranges)
    n = R.refshim_n_nodes(h)
    # lambda classes per node from the lambda tree (same topology, node value = class id)
    lt = oracle.parse_newick(re.sub(r"\)(\d+)", r")c\1", re.sub(r"([(,])(\d+)", r"\1c\2", lam_tree)))
    cls = np.array([int(nm[1:]) if nm.startswith("c") and nm[1:].isdigit() else 1 for nm in lt.names])
    lam_node = np.array([lams[c - 1] for c in cls])
    R.refshim_set_rates(h, d(lam_node), d(np.full(n, -1.0)))
    em_path = os.path.join(REF_IT, "errormodel.txt")
    fd, td_ = C.c_int(), C.c_int()
    dim = R.refshim_read_errormodel(em_path.encode(), ranges[1], None, C.byref(fd), C.byref(td_))
    E = np.zeros((dim, dim))
    R.refshim_read_errormodel(em_path.encode(), ranges[1], d(E), C.byref(fd), C.byref(td_))
    for leaf in range(0, n, 2):
        R.refshim_set_errormodel(h, leaf, d(E), dim, fd.value, td_.value)
    R.refshim_reset_cache(h)
    sample = np.arange(0, len(counts4), 25)
    Rr = ranges[3] - ranges[2] + 1
    L = np.zeros((len(sample), Rr))
    for k, f in enumerate(sample):
        R.refshim_likelihoods(h, i(np.ascontiguousarray(counts4[f])), d(L[k]))
    R.refshim_session_free(h)
    out.update(t4_newick=np.array(nw4), t4_species=np.array(species4), t4_table=table4.astype(np.int16), t4_lambda_tree=np.array(lam_tree),
               t4_lambdas=np.array(lams), t4_score=np.array(sc4), t4_poisson_lambda=np.array(float(pois4[0])),
               t4_root_range=np.array(rootr4), t4_family_range=np.array(famr4), t4_lam_node=lam_node,
               t4_sample=sample, t4_L=L, t4_errormodel_text=np.array(open(em_path).read()))
    print("test4: score", sc4, "families", len(counts4), "sample", len(sample))
    np.savez_compressed(os.path.join(OUT, "integration.npz"), **out)
    print("written", os.path.join(OUT, "integration.npz"), os.path.getsize(os.path.join(OUT, "integration.npz")), "bytes")


if __name__ == "__main__":
    main()
