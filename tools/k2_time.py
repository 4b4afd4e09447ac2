"""Time K1/K2 of one objective evaluation on a synthetic table (CUDA events recorded by the library); prints one JSON line.
Environment: CAFE_BENCH_FAMILIES / TAXA / MAXSIZE / MU (as bench.py's experiment overrides), K2_STEPS."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cafe_b200 import gpu as cgpu, host as chost, synth

F = int(os.environ.get("CAFE_BENCH_FAMILIES", 50000)); T = int(os.environ.get("CAFE_BENCH_TAXA", 20))
MS = int(os.environ.get("CAFE_BENCH_MAXSIZE", 200)); MU = float(os.environ.get("CAFE_BENCH_MU", 0))
steps = int(os.environ.get("K2_STEPS", 30))
newick = synth.random_tree(T, 1)
counts, lam0 = synth.simulate_table(newick, F, MS, seed=10)
uniq, mult, first = synth.dedup(counts)
tree = chost.parse_tree(newick)
rg = chost.init_family_size(MS)
ranges = (rg["min"], rg["max"], rg["root_min"], rg["root_max"])
R = ranges[3] - ranges[2] + 1
prior = chost.prior_poisson(ranges[2], 8.0, 1000)[:R]
n = tree.n_nodes
g = cgpu.CafeGpu(0)
g.set_tree(tree.left, tree.right, tree.branchlength); g.set_ranges(*ranges)
g.set_lnc_table(chost.lnc_table(max(ranges[1], ranges[3]))); g.set_families(uniq, mult, first); g.set_prior(prior)
mu = lambda lam: np.full(n, MU * lam) if MU > 0 else np.full(n, -1.0)
for k in range(5):
    lam = lam0 * (1 + 0.002 * k); s, z = g.objective(np.full(n, lam), mu(lam))
g.enable_timing(True)
for k in range(steps):
    lam = lam0 * (1 + 0.002 * (k % 40)); s, z = g.objective(np.full(n, lam), mu(lam))
k1, k2 = g.timing_collect()
fl = g.score_flops()
print(json.dumps({"k2_ms": float(np.mean(k2)), "k2_min_ms": float(np.min(k2)), "k1_ms": float(np.mean(k1)), "tflops": float(fl / (float(np.mean(k2)) * 1e-3) * 1e-12),
                  "score": float(s), "F": int(len(uniq)), "W": ranges[1] + 1}))
