"""Time the branch-stretch likelihood-ratio test (cafe_gpu_likelihood_ratio_test) at a BASELINE configs[1]-like shape on one GPU.
   python tools/time_lrt.py [n_taxa] [max_size] [n_families]"""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
from cafe_b200 import gpu as cgpu, host as chost, synth

n_taxa = int(sys.argv[1]) if len(sys.argv) > 1 else 20
max_size = int(sys.argv[2]) if len(sys.argv) > 2 else 200
F = int(sys.argv[3]) if len(sys.argv) > 3 else 50000
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
g.score()
l0 = g.launch_count()
t0 = time.perf_counter()
base, best, steps = g.likelihood_ratio_test()
t1 = time.perf_counter()
evals = int(steps.max(axis=1).sum() + (n - 1))  # batched evaluations: per branch (max steps + the one that stops everyone)
print(json.dumps({"lrt_seconds": t1 - t0, "n_taxa": n_taxa, "max_size": max_size, "families": int(len(uniq)), "branches": n - 1,
                  "batched_evaluations": evals, "ms_per_evaluation": 1e3 * (t1 - t0) / evals,
                  "family_branch_tests_per_s": len(uniq) * (n - 1) / (t1 - t0), "max_steps": int(steps.max()),
                  "mean_steps": float(steps.mean()), "launches": g.launch_count() - l0}))
