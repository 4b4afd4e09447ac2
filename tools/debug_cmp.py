"""Debug: per-family results of the default K2 against the first-generation kernel on the configs[1] table."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_parity import _config2_problem, _family_terms
nw, counts, lam0 = _config2_problem()
p, s, fz, lp, ml, am = _family_terms(nw, counts, lam0)
_, s1, fz1, lp1, ml1, am1 = _family_terms(nw, counts, lam0, env={"CAFE_GPU_FUSED_V1": "1"})
bad = np.nonzero((ml != ml1) | (am != am1))[0]
print("score", s, s1, "fz", fz, fz1, "bad", len(bad))
print("bad idx", bad[:40])
for i in bad[:12]:
    print(i, "blk", i // 8, "row", i % 8, "ml", ml[i], ml1[i], "am", am[i], am1[i], "lp", lp[i], lp1[i], "counts", counts[i].tolist())
_, s2, fz2, lp2, ml2, am2 = _family_terms(nw, counts, lam0)
print("repeat: identical", np.array_equal(ml, ml2), np.array_equal(lp, lp2))
