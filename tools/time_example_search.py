"""BASELINE configs[0]: the single-lambda search on example/example_data.tab through the host mirror (`lambda -s`), timed around the
command (device context already created by a first fixed-lambda command).  The table comes from tests/golden/example.npz.
   python tools/time_example_search.py"""
import json, os, sys, tempfile, time
import numpy as np
sys.path.insert(0, ".")
from cafe_b200 import host as chost

z = np.load("tests/golden/example.npz")
with tempfile.TemporaryDirectory() as td:
    p = os.path.join(td, "example_data.tab")
    with open(p, "w") as f:
        f.write("\t".join(["FAMILYDESC", "FAMILY"] + [str(s) for s in z["species_leaf_order"]]) + "\n")
        for i, r in zip(z["ids"], z["counts"]):
            f.write("\t".join(["d", str(i)] + [str(x) for x in r]) + "\n")
    s = chost.Session(quiet=True)
    for c in ("seed 10", "load -i %s -t 1" % p, "tree " + str(z["newick"]), "lambda -l 0.005", "seed 10"):
        assert s.command(c) == 0
    calls0 = s.objective_calls()
    t0 = time.perf_counter()
    assert s.command("lambda -s") == 0
    t1 = time.perf_counter()
    lam = s.parameters()[0]
    n = s.objective_calls() - calls0
    s.close()
print(json.dumps({"tool": "tools/time_example_search.py", "workload": "BASELINE configs[0]: example_data.tab (59 families, 5 taxa), lambda -s",
                  "search_seconds": t1 - t0, "objective_calls": n, "ms_per_objective_call": 1e3 * (t1 - t0) / max(n, 1), "lambda_hat": lam,
                  "reference_lambda_hat": float(z["search_lambda"]), "reference_cpu_seconds_whole_script_this_container": 0.72}))
