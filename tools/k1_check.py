"""K1: the anchored term recurrence (k_bd_matrix_rec, default) against the term-by-term kernel (CAFE_GPU_K1_EXACT=1) on the
bench shapes: largest relative difference over the entries > 1e-300, and both kernels' time per matrix build (CUDA events
recorded by the library).  One JSON line per shape."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_data
from cafe_b200 import gpu as cgpu, host as chost


def run(name, newick, maxsize, lam, mu_ratio):
    tree = chost.parse_tree(newick)
    rg = chost.init_family_size(maxsize)
    ranges = (rg["min"], rg["max"], rg["root_min"], rg["root_max"])
    maxfs = max(ranges[1], ranges[3])
    n = tree.n_nodes
    out = {"shape": name, "S": maxfs + 1}
    mats = {}
    for mode in ("rec", "exact"):
        if mode == "exact":
            os.environ["CAFE_GPU_K1_EXACT"] = "1"
        else:
            os.environ.pop("CAFE_GPU_K1_EXACT", None)
        g = cgpu.CafeGpu(0)
        g.set_tree(tree.left, tree.right, tree.branchlength); g.set_ranges(*ranges)
        g.set_lnc_table(chost.lnc_table(maxfs))
        n_leaves = (n + 1) // 2
        cnt = np.full((8, n_leaves), 3, dtype=np.int32)   # a token table: timing_collect pairs K1 with a K2
        g.set_families(cnt, np.ones(8, dtype=np.int32), np.arange(8, dtype=np.int32))
        g.set_prior(chost.prior_poisson(ranges[2], 8.0, 1000)[:ranges[3] - ranges[2] + 1])
        mu = np.full(n, mu_ratio * lam) if mu_ratio > 0 else np.full(n, -1.0)
        for _ in range(3):
            g.objective(np.full(n, lam), mu)
        g.enable_timing(True)
        for k in range(10):
            g.objective(np.full(n, lam * (1 + 1e-3 * k)), mu * (1 + 1e-3 * k) if mu_ratio > 0 else mu)
        k1, _ = g.timing_collect()
        out[f"k1_ms_{mode}"] = float(np.mean(k1))
        g.set_rates(np.full(n, lam), mu); g.build_matrices()
        seen = {}
        for v in range(n):
            if v != tree.root:
                seen.setdefault(int(tree.branchlength[v]), v)
        out["keys"] = len(seen)
        mats[mode] = [g.get_matrix(v) for _, v in sorted(seen.items())][:: max(1, len(seen) // 12)]
        g.close()
    os.environ.pop("CAFE_GPU_K1_EXACT", None)
    worst, where = 0.0, None
    for a, b in zip(mats["rec"], mats["exact"]):
        big = b > 1e-300
        rel = np.abs(a - b)[big] / b[big]
        if rel.size and rel.max() > worst:
            worst = float(rel.max())
        small = ~big
        assert np.abs(a - b)[small].max() < 1e-299 if small.any() else True
    out["max_rel_diff"] = worst
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    nw1 = bench_data.config_tree("configs[1]")
    run("configs[1] (20 taxa, max 200, lambda)", nw1, 200, bench_data.default_lambda(nw1), 0)
    nw2 = bench_data.config_tree("configs[2]")
    run("configs[2] (50 taxa, max 400, lambda/mu)", nw2, 400, bench_data.default_lambda(nw2), 0.8)
    run("configs[2] tree, lambda only", nw2, 400, bench_data.default_lambda(nw2), 0)
    run("S = 1001, lambda/mu", "((a:40,b:25):40,(c:11,d:63):17)", 800, 0.0015, 0.7)
    run("S = 1001, lambda", "((a:40,b:25):40,(c:11,d:63):17)", 800, 0.004, 0)
    run("small lambda t (steep ratios)", "((a:3,b:2):1,(c:1,d:5):2)", 400, 2e-5, 0.5)
