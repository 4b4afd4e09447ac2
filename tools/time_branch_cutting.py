"""Time `cafe_branch_cutting` (host mirror + cafe_gpu_cut_pvalues) at the BASELINE configs[1] shape: 50 k families x 20 taxa, max
size 200, 1000 random samples, the 5 % of the families with the smallest family-wide p-values tested on all 38 branches.
   python tools/time_branch_cutting.py [n_families] [n_taxa] [max_size] [n_samples]"""
import json, os, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cafe_b200 import host as chost, synth

F = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 20
MS = int(sys.argv[3]) if len(sys.argv) > 3 else 200
N = int(sys.argv[4]) if len(sys.argv) > 4 else 1000
nw = synth.random_tree(T, 1)
counts, lam0 = synth.simulate_table(nw, F, MS, seed=10)
tree = chost.parse_tree(nw)
species = [tree.names[i] for i in range(0, tree.n_nodes, 2)]
d = tempfile.mkdtemp()
path = os.path.join(d, "fam.tab")
with open(path, "w") as f:
    f.write("\t".join(["Desc", "Family ID"] + species) + "\n")
    for i, r in enumerate(counts):
        f.write("\t".join(["d", "F%d" % i] + [str(int(x)) for x in r]) + "\n")
s = chost.Session(quiet=True)
assert s.command("load -i %s -t 8 -r %d -p 0.05" % (path, N)) == 0      # 8 threads: device RNG instead of the rand() replay
assert s.command("tree " + nw) == 0
assert s.command("lambda -l %.10g" % lam0) == 0
rs = np.random.RandomState(2)
maxp = rs.uniform(0.051, 1.0, len(counts))
tested = rs.choice(len(counts), len(counts) // 20, replace=False)
maxp[tested] = rs.uniform(0, 0.05, len(tested))
s.set_max_pvalues(maxp)
t0 = time.perf_counter()
cut = s.branch_cutting(N)
dt = time.perf_counter() - t0
ok = cut[:, tested]
print(json.dumps({"families": int(len(counts)), "tested_families": int(len(tested)), "taxa": T, "branches": int(cut.shape[0] - 1), "max_size": MS,
                  "random_samples": N, "seconds": dt, "branch_family_pvalues_per_s": len(tested) * (cut.shape[0] - 1) / dt,
                  "mean_cut_pvalue_tested": float(ok[ok >= 0].mean()), "untested_all_minus_one": bool((np.delete(cut, tested, axis=1) == -1).all())}))
