"""The branch-stretch likelihood-ratio test across ranks (launch under torchrun): ONE table of F families split over the ranks
(sharding.likelihood_ratio_test_sharded), strong scaling; the checksum of the gathered result must not depend on the rank count.
   torchrun --nproc-per-node N tools/run_lrt_sharded.py [n_taxa] [max_size] [families_total]"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
import torch.distributed as dist
from cafe_b200 import gpu as cgpu, host as chost, sharding, synth

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n_taxa = int(sys.argv[1]) if len(sys.argv) > 1 else 20
max_size = int(sys.argv[2]) if len(sys.argv) > 2 else 200
F = int(sys.argv[3]) if len(sys.argv) > 3 else 50000
nw = synth.random_tree(n_taxa, 1)
counts, lam0 = synth.simulate_table(nw, F, max_size, seed=10, device=local)   # the same table on every rank
uniq, mult, first = synth.dedup(counts)
lo, hi = sharding.shard_bounds(len(uniq), world, rank)
tree = chost.parse_tree(nw)
rg = chost.init_family_size(max_size)
ranges = (rg["min"], rg["max"], rg["root_min"], rg["root_max"])
R = ranges[3] - ranges[2] + 1
g = cgpu.CafeGpu(local)
g.set_tree(tree.left, tree.right, tree.branchlength)
g.set_ranges(*ranges)
g.set_lnc_table(chost.lnc_table(max(ranges[1], ranges[3])))
g.set_families(uniq[lo:hi], mult[lo:hi], first[lo:hi])
g.set_prior(chost.prior_poisson(ranges[2], 8.0, 1000)[:R])
n = tree.n_nodes
g.set_rates(np.full(n, lam0), np.full(n, -1.0))
g.build_matrices()
g.score()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
base, best, steps = sharding.likelihood_ratio_test_sharded(g, np.ones(hi - lo, dtype=np.uint8), len(uniq), rank, world)
torch.cuda.synchronize()
t1 = time.perf_counter()
tt = torch.tensor([t1 - t0], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"tool": "tools/run_lrt_sharded.py", "n_gpus": world, "n_taxa": n_taxa, "max_size": max_size, "families_total": int(len(uniq)),
                      "branches": n - 1, "lrt_seconds_max_over_ranks": tt.item(), "includes": "the all-gather of the result rows to every rank",
                      "checksum_log_best": float(np.log(best[best > 0]).sum()), "steps_total": int(steps.sum())}))
if world > 1:
    dist.destroy_process_group()
