"""Time K4 (conditional distribution), K5 (family p-values) and the Viterbi reconstruction at a BASELINE configs[4]-like shape on one GPU.
   python tools/time_cd_pvalue.py [n_taxa] [max_size] [n_samples] [n_families]"""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
from cafe_b200 import gpu as cgpu, host as chost, synth

n_taxa = int(sys.argv[1]) if len(sys.argv) > 1 else 50
max_size = int(sys.argv[2]) if len(sys.argv) > 2 else 400
n_samples = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
F = int(sys.argv[4]) if len(sys.argv) > 4 else 25000
nw = synth.random_tree(n_taxa, 1)
counts, lam0 = synth.simulate_table(nw, F, max_size, seed=10)
tree = chost.parse_tree(nw)
rg = chost.init_family_size(max_size)
ranges = (rg["min"], rg["max"], rg["root_min"], rg["root_max"])
R = ranges[3] - ranges[2] + 1
g = cgpu.CafeGpu()
g.set_tree(tree.left, tree.right, tree.branchlength)
g.set_ranges(*ranges)
g.set_lnc_table(chost.lnc_table(max(ranges[1], ranges[3])))
uniq, mult, first = synth.dedup(counts)
g.set_families(uniq, mult, first)
g.set_prior(chost.prior_poisson(ranges[2], 8.0, 1000)[:R])
n = tree.n_nodes
g.set_rates(np.full(n, lam0), np.full(n, -1.0))
g.build_matrices()
t0 = time.perf_counter()
cd = g.conditional_distribution(n_samples, seed=7)
t1 = time.perf_counter()
pv = g.pvalues(cd)
t2 = time.perf_counter()
t3 = time.perf_counter()
sizes, ml = g.viterbi()
t4a = time.perf_counter()
g.viterbi_report()
t4b = time.perf_counter()
g.viterbi_report()
t4c = time.perf_counter()
t4 = t4a
print(json.dumps({"viterbi_report_seconds": t4c - t4b, "viterbi_report_first_call_seconds": t4b - t4a, "viterbi_seconds": t4 - t3, "viterbi_families_per_s": len(uniq) / (t4 - t3), "n_taxa": n_taxa, "max_size": max_size, "R": R, "n_samples": n_samples, "families": int(len(uniq)),
                  "cd_seconds": t1 - t0, "simulated_prunings_per_s": R * n_samples / (t1 - t0),
                  "pvalue_seconds": t2 - t1, "family_pvalues_per_s": len(uniq) / (t2 - t1),
                  "pvalue_mean": float(np.mean(pv)), "launches": g.launch_count()}))
