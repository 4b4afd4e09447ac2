import os, sys, time
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np
from cafe_b200 import synth
from util import Problem
nw = synth.random_tree(20, 1)
counts, lam0 = synth.simulate_table(nw, 170000, 200, seed=10)
res = {}
for name, env in (("v2", None), ("v1", "CAFE_GPU_FUSED_V1")):
    if env: os.environ[env] = "1"
    p = Problem(nw, counts, lam0, prior_lambda=8.0)
    g = p.make_gpu()
    t0 = time.time(); s, fz = g.score(); t1 = time.time()
    lp, ml, am = g.family_results()
    g.close()
    if env: os.environ.pop(env)
    res[name] = (s, fz, lp, ml, am)
    print(name, s, fz, round(t1 - t0, 4))
a, b = res["v2"], res["v1"]
print("equal maxlik", np.array_equal(a[3], b[3]), "equal argmax", np.array_equal(a[4], b[4]), "max |dlogpost|", np.abs(a[2] - b[2]).max())
