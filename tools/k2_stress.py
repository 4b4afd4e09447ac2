"""Race hunt: evaluate the configs[1] table N times with the default K2 and count the families whose results differ from the
first-generation kernel's (bit-identical arithmetic).  CAFE_GPU_LIB selects an A/B build."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_parity import _config2_problem, Problem
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
nw, counts, lam0 = _config2_problem()
p = Problem(nw, counts, lam0, prior_lambda=8.0)
os.environ["CAFE_GPU_FUSED_V1"] = "1"
g = p.make_gpu(); g.score(); lp1, ml1, am1 = g.family_results(); g.close()
del os.environ["CAFE_GPU_FUSED_V1"]
fresh = os.environ.get("K2_STRESS_FRESH") == "1"   # a new context (cold scratch, TLB, L2) per evaluation
g = p.make_gpu()
bad_runs, bad_total, first = 0, 0, None
for it in range(n):
    if fresh and it > 0:
        g.close(); g = p.make_gpu()
    s, fz = g.score()
    lp, ml, am = g.family_results()
    bad = np.nonzero((ml != ml1) | (am != am1))[0]
    if len(bad):
        bad_runs += 1; bad_total += len(bad)
        if first is None: first = bad[:24].tolist()
print({"lib": os.environ.get("CAFE_GPU_LIB", "default"), "runs": n, "fresh": fresh, "bad_runs": bad_runs, "bad_families": bad_total, "first": first})
