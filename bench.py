#!/usr/bin/env python
"""bench.py — gene-family log-likelihoods/sec per lambda-evaluation (BASELINE.json metric).

One "step" = one objective evaluation of CAFE's lambda search (seam B1): K1 transition-matrix build +
K2 batched pruning over all families + K3 score reduction (+ one 2-double collective when N > 1).

Workload at every N: BASELINE.json configs[1] per GPU — 50 k synthetic families simulated from the
birth–death model on a 20-taxon integer-branch-length tree, observed max size 200 (=> W=251, R=250,
S=251), single lambda.  Weak scaling: each rank holds its own 50 k families, every rank builds all
matrices, the only exchange is the reduction of {partial score, first zero family}.

  python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm (CUDA through the C-ABI)
  python bench.py --impl reference ...                           # the reference's CPU path (oracle/_ref)

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "gene-family log-likelihoods/sec per lambda-eval"
UNIT = "families/s"
# The headline workload is BASELINE configs[1].  The CAFE_BENCH_* overrides exist for kernel experiments and for informational
# lines on the other BASELINE shapes (e.g. configs[2] per GPU at 8 GPUs: CAFE_BENCH_FAMILIES=25000 CAFE_BENCH_TAXA=50
# CAFE_BENCH_MAXSIZE=400 CAFE_BENCH_MU=0.8); the JSON line always names the workload it actually ran.
N_TAXA = int(os.environ.get("CAFE_BENCH_TAXA", 20))
FAMILIES_PER_GPU = int(os.environ.get("CAFE_BENCH_FAMILIES", 50000))
MAX_SIZE = int(os.environ.get("CAFE_BENCH_MAXSIZE", 200))
MU_RATIO = float(os.environ.get("CAFE_BENCH_MU", 0))  # > 0: lambdamu mode with mu = ratio * lambda
TREE_SEED = 1
IS_HEADLINE = (N_TAXA, FAMILIES_PER_GPU, MAX_SIZE, MU_RATIO) == (20, 50000, 200, 0)
WORKLOAD_NAME = "BASELINE configs[1]" if IS_HEADLINE else "experiment override, not the headline workload"


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self):
        return len(self.lines)

    def stop(self, lo=0, hi=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines[lo:hi]:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
            except ValueError:
                continue
            for nm, val in zip(names, p[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def workload(rank):
    """(newick, counts[F][n_leaves], lam0) for this rank; tree shared, families seeded per rank."""
    from cafe_b200 import synth
    newick = synth.random_tree(N_TAXA, TREE_SEED)
    counts, lam0 = synth.simulate_table(newick, FAMILIES_PER_GPU, MAX_SIZE, seed=10 + rank)
    return newick, counts, lam0


def lambda_schedule(lam0, k):
    # a different lambda every step, as successive Nelder–Mead vertices would be
    return lam0 * (1.0 + 0.002 * (k % 40))


# =====================================================================================================
# reference arm: the reference's own CPU implementation (oracle/_ref), bounded sample per step
# =====================================================================================================
def cpu_reference_eval(newick, counts_sample, lam, ranges, prior, n_total):
    """Time one objective evaluation of the compiled reference on a family sample.
    Returns dict(t_matrices, t_sample, value) with value = n_total / (t_matrices + n_total/len(sample)*t_sample)."""
    import ctypes as C
    import oracle
    R = oracle.ref()
    kind = "reference"
    dp = C.POINTER(C.c_double)
    ip = C.POINTER(C.c_int)
    if R is None:  # reference not compiled here: fall back to the oracle port (still a CPU baseline)
        kind = "port"
        t = oracle.parse_newick(newick)
        n = t.n_nodes
        t0 = time.perf_counter()
        mats = oracle.node_matrices(t, [lam] * n, [-1.0] * n, max(ranges[1], ranges[3]))
        t1 = time.perf_counter()
        oracle.score(t, mats, counts_sample, ranges, prior)
        t2 = time.perf_counter()
    else:
        h = R.refshim_session_new(newick.encode(), *ranges)
        n = R.refshim_n_nodes(h)
        lam_a = np.full(n, lam)
        mu_a = np.full(n, -1.0)
        R.refshim_set_rates(h, lam_a.ctypes.data_as(dp), mu_a.ctypes.data_as(dp))
        cs = np.ascontiguousarray(counts_sample, dtype=np.int32)
        R.refshim_set_families(h, len(cs), cs.ctypes.data_as(ip), 0)
        pr = np.zeros(1000)
        pr[: len(prior)] = prior
        t0 = time.perf_counter()
        R.refshim_reset_cache(h)  # reset_birthdeath_cache: all matrices, cafe_main.c:319
        t1 = time.perf_counter()
        threw = C.c_int(0)
        R.refshim_get_posterior(h, pr.ctypes.data_as(dp), C.byref(threw), None, 0)  # get_posterior, lambda.cpp:691
        t2 = time.perf_counter()
        R.refshim_session_free(h)
    t_mat, t_fam = t1 - t0, t2 - t1
    per_eval = t_mat + n_total / len(counts_sample) * t_fam
    return {"t_matrices": t_mat, "t_sample": t_fam, "value": n_total / per_eval, "kind": kind}


def run_reference(args, rank, world):
    if rank != 0:
        return
    from cafe_b200 import host as chost
    import oracle
    from cafe_b200 import synth
    newick = synth.random_tree(N_TAXA, TREE_SEED)
    # family sample of the same shape, drawn on the CPU with the oracle's matrices (no GPU on this arm)
    ot = oracle.parse_newick(newick)
    lam0 = 0.25 / synth.tree_depth(chost.parse_tree(newick))
    roots = np.r_[1 + np.random.RandomState(3).poisson(8.0, 850), np.random.RandomState(4).randint(1, MAX_SIZE + 1, 150)]
    counts = oracle.simulate_families(ot, [lam0] * ot.n_nodes, [-1.0] * ot.n_nodes, 250,
                                      max(2000, 2 * args.ref_sample), roots, 10)
    counts = counts[counts.max(axis=1) <= MAX_SIZE]
    rg = chost.init_family_size(MAX_SIZE)
    ranges = (rg["min"], rg["max"], rg["root_min"], rg["root_max"])
    prior = chost.prior_poisson(ranges[2], 8.0, 1000)[: ranges[3] - ranges[2] + 1]
    sample_n = args.ref_sample
    n_total = FAMILIES_PER_GPU * args.gpus
    vals, last = [], None
    for k in range(args.warmup + args.steps):
        lo = (k * sample_n) % max(1, len(counts) - sample_n)
        r = cpu_reference_eval(newick, counts[lo:lo + sample_n], lambda_schedule(lam0, k), ranges, prior, n_total)
        if k >= args.warmup:
            vals.append(r["value"])
        last = r
    value = float(np.mean(vals))
    sample = (f"{sample_n} of {n_total} families per step + the full matrix build; per-eval time = t_matrices + "
              f"(F/sample)*t_sample (the family loop is exactly linear, SURVEY.md 8d)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * n_total / value, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic (simulated from the birth-death model)",
        "config": {"workload": f"{FAMILIES_PER_GPU} families x {N_TAXA} taxa, max size {MAX_SIZE}, single lambda ({WORKLOAD_NAME}) per GPU-equivalent",
                   "families_total": n_total, "W": ranges[1] + 1, "R": ranges[3] - ranges[2] + 1},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": last["kind"], "sample": sample,
                         "note": "the reference's lambda search is single-threaded over families (lambda.cpp:698-722)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# =====================================================================================================
# our arm
# =====================================================================================================
def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from cafe_b200 import gpu as cgpu
    from cafe_b200 import host as chost
    from cafe_b200 import sharding, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    newick, counts, lam0 = workload(rank)
    uniq, mult, first = synth.dedup(counts)
    first = first + rank * FAMILIES_PER_GPU  # global list index of each pattern's first family
    tree = chost.parse_tree(newick)
    rg = chost.init_family_size(MAX_SIZE)
    ranges = (rg["min"], rg["max"], rg["root_min"], rg["root_max"])
    R = ranges[3] - ranges[2] + 1
    prior = chost.prior_poisson(ranges[2], 8.0, 1000)[:R]
    n = tree.n_nodes
    mu_node = np.full(n, -1.0)

    def mu_of(lam):
        return np.full(n, MU_RATIO * lam) if MU_RATIO > 0 else mu_node

    g = cgpu.CafeGpu(local_rank)
    # everything (our kernels, the NCCL collective, the timing events) runs on one explicit torch stream
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    g.set_stream(stream.cuda_stream)
    g.set_tree(tree.left, tree.right, tree.branchlength)
    g.set_ranges(*ranges)
    g.set_lnc_table(chost.lnc_table(max(ranges[1], ranges[3])))
    g.set_families(uniq, mult, first)
    g.set_prior(prior)
    out2 = torch.zeros(2, dtype=torch.float64, device=dev)

    shard_k1 = world > 1 and os.environ.get("CAFE_BENCH_NO_K1_SHARD") is None
    if shard_k1:
        g.set_key_shard(rank, world)  # every rank builds 1/world of the matrices, NCCL all-gathers them (sharding.objective_sharded)

    def step_device(k):
        lam_k = lambda_schedule(lam0, k)
        if shard_k1:
            return sharding.objective_sharded(g, np.full(n, lam_k), mu_of(lam_k), out2, rank, world, dev)
        g.objective_device(np.full(n, lam_k), mu_of(lam_k), out2.data_ptr())
        return sharding.reduce_score(out2)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # ---- warm-up ----
    for k in range(args.warmup):
        s, z = step_device(k)
    barrier()
    score_chk, _ = sharding.finish_score(s, z) if args.warmup else (0.0, -1)

    # ---- device-timed region: K steps, inputs resident in HBM ----
    g.enable_timing(True)
    g.reset_launch_count()
    if rank == 0:  # wait for the first nvidia-smi sample so that the timed region is covered
        t_wait = time.time()
        while sampler.mark() == 0 and time.time() - t_wait < 3.0:
            time.sleep(0.01)
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    barrier()
    mark0 = sampler.mark()
    e0.record(stream)
    for k in range(args.steps):
        s, z = step_device(args.warmup + k)
    e1.record(stream)
    barrier()
    mark1 = sampler.mark() + 1
    ms_total = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    clocks = sampler.stop(max(0, mark0 - 1), mark1) if rank == 0 else None
    launches = g.launch_count() + (args.steps * (3 if shard_k1 else 1) if world > 1 else 0)  # + the NCCL all-gathers of every step
    k1_ms, k2_ms = g.timing_collect()
    g.enable_timing(False)
    ms_per_step = float(ms_total.item()) / args.steps
    families_total = FAMILIES_PER_GPU * world
    value = families_total / (ms_per_step * 1e-3)
    last_score, last_zero = sharding.finish_score(s, z)

    # ---- end-to-end: the reference-facing call with HOST buffers in and out, every step ----
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        lam_node = np.full(n, lambda_schedule(lam0, args.warmup + k))
        if world == 1:
            sc, fz = g.objective(lam_node, mu_of(lam_node[0]))  # host lambda array in, host score out (sync inside)
        else:
            if shard_k1:
                s2, z2 = sharding.objective_sharded(g, lam_node, mu_of(lam_node[0]), out2, rank, world, dev)
            else:
                g.objective_device(lam_node, mu_of(lam_node[0]), out2.data_ptr())
                s2, z2 = sharding.reduce_score(out2)
            sc, fz = sharding.finish_score(s2.cpu(), z2.cpu())
    torch.cuda.synchronize()
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = families_total / (float(t_e2e.item()) / args.steps)
    h2d = g.num_keys() * 48  # BdKeyParams per distinct (int t, lambda, mu) key; rates are de-duplicated on the host
    d2h = 16

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- fp64 roofline denominator: cuBLAS DGEMM on this GPU, this run ----
    a = torch.randn(6144, 6144, dtype=torch.float64, device=dev)
    b = torch.randn(6144, 6144, dtype=torch.float64, device=dev)
    c = torch.empty_like(a)
    for _ in range(2):
        torch.matmul(a, b, out=c)
    best = 1e9
    for _ in range(4):
        x0 = torch.cuda.Event(enable_timing=True)
        x1 = torch.cuda.Event(enable_timing=True)
        x0.record()
        torch.matmul(a, b, out=c)
        x1.record()
        torch.cuda.synchronize()
        best = min(best, x0.elapsed_time(x1))
    dgemm_tflops = 2 * 6144 ** 3 / best * 1e-9
    del a, b, c

    flops = g.score_flops()  # algorithmic: internal edges only, 2*W*W (2*R*W under the root) per unique family
    k2 = float(np.mean(k2_ms)) if len(k2_ms) else float("nan")
    k1 = float(np.mean(k1_ms)) if len(k1_ms) else float("nan")
    achieved = flops / (k2 * 1e-3) * 1e-12
    traffic = None
    tp = os.path.join(ROOT, "profiles", "k2_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_step")
        except Exception:
            traffic = None

    # ---- CPU baseline on this box's host cores (bounded sample) ----
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_eval(newick, counts[: args.cpu_sample], lam0, ranges, prior, FAMILIES_PER_GPU)
        cpu = {"value": r["value"], "unit": UNIT, "cores": 1, "kind": r["kind"],
               "sample": f"{args.cpu_sample} of {FAMILIES_PER_GPU} families + the full matrix build "
                         f"({r['t_matrices']:.2f}s matrices, {r['t_sample']:.2f}s sample), extrapolated linearly in F",
               "host_cpus": os.cpu_count()}

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic (simulated from the birth-death model at lambda0=0.25/depth)",
        "config": {
            "workload": f"{FAMILIES_PER_GPU} families x {N_TAXA} taxa per GPU, max size {MAX_SIZE}, " + ("lambda and mu" if MU_RATIO > 0 else "single lambda") + f" ({WORKLOAD_NAME})",
            "families_total": families_total, "unique_patterns_rank0": int(len(uniq)), "W": ranges[1] + 1, "R": R,
            "S": max(ranges[1], ranges[3]) + 1, "keys": g.num_keys(), "parallelism": f"families sharded x{world}" + (f", matrix build sharded x{world} + 2 NCCL all-gathers" if shard_k1 else ""),
            "l2": "no explicit flush: the node-vector scratch of K2 (0.9 GB, 4.9 GB of DRAM traffic per launch) exceeds the 126 MB L2 and the matrices are rewritten every step",
            "last_score": last_score, "k1_ms": k1, "k2_ms": k2,
        },
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "what": "cafe_gpu_objective(host lambda/mu arrays) -> host score, per step; the family table is session state set once, as after the reference's `load`"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": dgemm_tflops, "unit": "TFLOP/s",
                     "frac": achieved / dgemm_tflops, "traffic": traffic,
                     "kernel": "K2 batched pruning (DMMA.8x8x4 fp64)", "flops_per_step": flops,
                     "peak_source": "cuBLAS DGEMM 6144^3 fp64 measured in this run (MEASURED_PEAKS.json has no fp64 figure; "
                                    "tcgen05 has no fp64 kind, the fp64 tensor pipe is DMMA)",
                     "hbm_peak_gbs_measured": peaks.get("hbm_gbs")},
    }
    if cpu:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    g.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=1500)
    ap.add_argument("--ref-sample", type=int, default=300)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank, local_rank, world = env_int("RANK", 0), env_int("LOCAL_RANK", 0), env_int("WORLD_SIZE", 1)
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
