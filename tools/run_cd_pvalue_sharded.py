"""BASELINE configs[4] across ranks (launch under torchrun): the conditional distribution with its root-size rows split over the
ranks (sharding.conditional_distribution_sharded), then the family p-values of each rank's own families.
   torchrun --nproc-per-node N tools/run_cd_pvalue_sharded.py [n_taxa] [max_size] [n_samples] [families_per_rank]"""
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
n_taxa = int(sys.argv[1]) if len(sys.argv) > 1 else 50
max_size = int(sys.argv[2]) if len(sys.argv) > 2 else 400
n_samples = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
F = int(sys.argv[4]) if len(sys.argv) > 4 else 25000
nw = synth.random_tree(n_taxa, 1)
counts, lam0 = synth.simulate_table(nw, F, max_size, seed=10 + rank, device=local)
tree = chost.parse_tree(nw)
rg = chost.init_family_size(max_size)
ranges = (rg["min"], rg["max"], rg["root_min"], rg["root_max"])
R = ranges[3] - ranges[2] + 1
g = cgpu.CafeGpu(local)
g.set_tree(tree.left, tree.right, tree.branchlength)
g.set_ranges(*ranges)
g.set_lnc_table(chost.lnc_table(max(ranges[1], ranges[3])))
uniq, mult, first = synth.dedup(counts)
g.set_families(uniq, mult, first)
g.set_prior(chost.prior_poisson(ranges[2], 8.0, 1000)[:R])
n = tree.n_nodes
g.set_rates(np.full(n, lam0), np.full(n, -1.0))
g.build_matrices()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
cd = sharding.conditional_distribution_sharded(g, n_samples, 7, rank, world)
t1 = time.perf_counter()
pv = g.pvalues(cd)
torch.cuda.synchronize()
t2 = time.perf_counter()
tt = torch.tensor([t1 - t0, t2 - t1], dtype=torch.float64, device="cuda")
chk = torch.tensor([float(cd.sum())], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    lo = chk.clone(); hi = chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    same = bool(lo.item() == hi.item())
else:
    same = True
if rank == 0:
    print(json.dumps({"n_gpus": world, "n_taxa": n_taxa, "max_size": max_size, "R": R, "n_samples": n_samples, "families_per_rank": int(len(uniq)),
                      "cd_seconds_max_over_ranks": tt[0].item(), "pvalue_seconds_max_over_ranks": tt[1].item(),
                      "same_distribution_on_every_rank": same, "cd_checksum": chk.item(), "pvalue_mean_rank0": float(np.mean(pv))}))
if world > 1:
    dist.destroy_process_group()
