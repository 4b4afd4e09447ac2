#!/usr/bin/env python
"""bench.py — gene-family log-likelihoods/sec per lambda-evaluation (BASELINE.json metric).

One "step" = one objective evaluation of CAFE's lambda / lambda-mu search (seam B1, cafe/lambda.cpp:726-769,
cafe/lambdamu.cpp:323-367): K1 transition-matrix build + K2 batched pruning over all families + K3 score reduction and, on
N > 1 GPUs, the two NCCL exchange steps the C-ABI library runs itself (all-gather of the sharded matrices, reduction of the
score).

Headline workload at EVERY N: BASELINE.json configs[2] — the configuration the metric's "at 1/2/4/8 B200" is quoted on:
ONE fixed table of 200 k synthetic families on a 50-taxon integer-branch-length tree, max size 400 (=> W=481, R=500, S=501),
separate lambda and mu.  STRONG scaling: rank r of N holds the r-th contiguous 1/N of the table (N=1 holds all of it), every
rank builds 1/N of the 98 distinct matrices.  At N=1 the line also carries `configs`, with configs[1] (50 k x 20 taxa, single
lambda), configs[3] (100 k x 100 taxa, 4 lambda classes, error model) and configs[4] (conditional distribution + p-values)
measured the same way.

  python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm (CUDA through the C-ABI)
  python bench.py --impl reference ...                           # the reference's own CPU path (oracle/_ref), same table

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

import bench_data  # noqa: E402  (pure numpy: the same table for both arms)

METRIC = "gene-family log-likelihoods/sec per lambda-eval"
UNIT = "families/s"
HEADLINE = os.environ.get("CAFE_BENCH_CONFIG", "configs[2]")
DATA = "synthetic (simulated from the linear birth-death process at lambda0 = 0.25/depth, bench_data.py)"


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def workload_label(name):
    return bench_data.CONFIGS[name]["label"] + ", one fixed table split over the GPUs"


def lambda_schedule(lam0, k):
    # a different lambda every step, as successive Nelder–Mead vertices would be
    return lam0 * (1.0 + 0.002 * (k % 40))


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
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines[lo:hi]:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
                pw.append(float(p[3]))
            except ValueError:
                continue
            for nm, val in zip(names, p[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": float(max(pw)) if pw else None}


# =====================================================================================================
# reference arm: the reference's own CPU implementation (oracle/_ref), bounded sample per step.
# Imports nothing of the product (no cafe_b200): tree, ranges and prior come from the reference itself.
# =====================================================================================================
class ReferenceCpu:
    """The compiled reference (OpenMP build, Makefile.in:13) on one configuration: full matrix build with all host threads
    (`omp for` over the keys, cafe_tree.c:468) + get_posterior over a family sample (serial over families, lambda.cpp:698-722;
    two `omp task`s per node, cafe_tree.c:252-260 — the thread count that is fastest here is calibrated once)."""

    def __init__(self, name):
        import ctypes as C
        import oracle
        self.C = C
        self.kind = "reference"
        self.R = oracle.ref(openmp=True) or oracle.ref()
        self.name = name
        self.cfg = bench_data.CONFIGS[name]
        self.newick = bench_data.config_tree(name)
        self.lam0 = bench_data.default_lambda(self.newick)
        self.cores = os.cpu_count() or 1
        try:
            self.gomp = C.CDLL("libgomp.so.1") if oracle.ref(openmp=True) is not None else None
        except OSError:
            self.gomp = None
        self.family_threads = 1
        if self.R is None:
            # the reference was not compiled here: the oracle's C restatement of the same algorithm (1 core) stands in
            self.kind = "port"
            self.oracle = oracle
            self.tree = oracle.parse_newick(self.newick)
            a = [C.c_int() for _ in range(4)]
            oracle.lib().orc_init_family_size(self.cfg["max_size"], *[C.byref(x) for x in a])  # root_min, root_max, min, max
            self.ranges = (a[2].value, a[3].value, a[0].value, a[1].value)
            self.prior = bench_data.root_prior(self.cfg["max_size"], self.ranges[2], 1000)
            self.cores = 1
            return
        rg = (C.c_int * 4)()
        self.R.refshim_init_family_size(self.cfg["max_size"], rg)  # init_family_size, cafe_family.c:357-364
        self.ranges = (rg[2], rg[3], rg[0], rg[1])                 # min, max, root_min, root_max
        self.prior = bench_data.root_prior(self.cfg["max_size"], self.ranges[2], 1000)  # the tables' own root distribution

    def threads(self, n):
        if self.gomp is not None:
            self.gomp.omp_set_num_threads(int(n))

    def rates(self, n_nodes, lam):
        mu = self.cfg["mu_ratio"] * lam if self.cfg["mu_ratio"] > 0 else -1.0
        if self.name == "configs[3]":
            return lam * np.array([1.0, 1.3, 0.7, 1.1])[np.arange(n_nodes) % 4], np.full(n_nodes, mu)
        return np.full(n_nodes, lam), np.full(n_nodes, mu)

    def evaluate(self, counts_sample, lam, n_total):
        """One objective evaluation on a family sample; value = n_total / (t_matrices + n_total / len(sample) * t_sample)."""
        C = self.C
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        if self.kind == "port":
            t = self.tree
            la, mu = self.rates(t.n_nodes, lam)
            t0 = time.perf_counter()
            mats = self.oracle.node_matrices(t, list(la), list(mu), max(self.ranges[1], self.ranges[3]))
            t1 = time.perf_counter()
            score = self.oracle.score(t, mats, np.ascontiguousarray(counts_sample, dtype=np.int32), self.ranges,
                                      self.prior[: self.ranges[3] - self.ranges[2] + 1])["score"]
            t2 = time.perf_counter()
            t_mat, t_fam = t1 - t0, t2 - t1
            per_eval = t_mat + n_total / len(counts_sample) * t_fam
            return {"t_matrices": t_mat, "t_sample": t_fam, "value": n_total / per_eval, "score": score, "per_eval_s": per_eval}
        R = self.R
        h = R.refshim_session_new(self.newick.encode(), *self.ranges)
        n = R.refshim_n_nodes(h)
        la, mu = self.rates(n, lam)
        R.refshim_set_rates(h, la.ctypes.data_as(dp), mu.ctypes.data_as(dp))
        cs = np.ascontiguousarray(counts_sample, dtype=np.int32)
        R.refshim_set_families(h, len(cs), cs.ctypes.data_as(ip), 0)
        self.threads(self.cores)
        t0 = time.perf_counter()
        R.refshim_reset_cache(h)  # reset_birthdeath_cache: all matrices, cafe_main.c:319
        t1 = time.perf_counter()
        self.threads(self.family_threads)
        threw = C.c_int(0)
        score = R.refshim_get_posterior(h, self.prior.ctypes.data_as(dp), C.byref(threw), None, 0)  # get_posterior, lambda.cpp:691
        t2 = time.perf_counter()
        R.refshim_session_free(h)
        t_mat, t_fam = t1 - t0, t2 - t1
        per_eval = t_mat + n_total / len(cs) * t_fam
        return {"t_matrices": t_mat, "t_sample": t_fam, "value": n_total / per_eval, "score": score, "per_eval_s": per_eval}

    def cond_dist_sample(self, n_rows, n_samples):
        """The reference's pthreads conditional distribution (cafe_conditional_distribution, conditional_distribution.cpp:86-120:
        one thread per block of root sizes) on `n_rows` root sizes x `n_samples` draws with all host cores; draws per second."""
        C = self.C
        dp = C.POINTER(C.c_double)
        R = self.R
        rg = (self.ranges[0], self.ranges[1], self.ranges[2], self.ranges[2] + n_rows - 1)
        h = R.refshim_session_new(self.newick.encode(), *rg)
        n = R.refshim_n_nodes(h)
        la, mu = self.rates(n, self.lam0)
        R.refshim_set_rates(h, la.ctypes.data_as(dp), mu.ctypes.data_as(dp))
        self.threads(self.cores)
        R.refshim_reset_cache(h)
        out = np.zeros((n_rows, n_samples))
        R.refshim_srand(10)
        t0 = time.perf_counter()
        R.refshim_cond_dist(h, self.cores, n_samples, out.ctypes.data_as(dp))
        dt = time.perf_counter() - t0
        R.refshim_session_free(h)
        return n_rows * n_samples / dt, dt

    def calibrate(self, counts_sample, lam):
        """Pick the OpenMP thread count that makes the (serial-over-families) posterior loop fastest on this box."""
        if self.gomp is None or self.kind != "reference":
            return
        best = None
        for nt in sorted({1, 2, self.cores}):
            self.family_threads = nt
            r = self.evaluate(counts_sample, lam, len(counts_sample))
            if best is None or r["t_sample"] < best[0]:
                best = (r["t_sample"], nt)
        self.family_threads = best[1]


def run_reference(args, rank, world):
    if rank != 0:
        return
    name = HEADLINE
    ref = ReferenceCpu(name)
    counts = bench_data.config_chunk(name, 0)  # the first rows of the GPU arm's table
    n_total = ref.cfg["families"]
    sample_n = args.ref_sample
    ref.calibrate(counts[:max(4, sample_n // 4)], ref.lam0)
    vals, last = [], None
    for k in range(args.warmup + args.steps):
        lo = (k * sample_n) % max(1, len(counts) - sample_n)
        r = ref.evaluate(counts[lo:lo + sample_n], lambda_schedule(ref.lam0, k), n_total)
        if k >= args.warmup:
            vals.append(r["value"])
        last = r
    value = float(np.mean(vals))
    sample = (f"{sample_n} of {n_total} families per step (rows of chunk 0 of the same table) + the FULL matrix build; per-eval time = "
              f"t_matrices ({last['t_matrices']:.2f}s, omp for over keys, {ref.cores} threads) + (F/sample) * t_sample "
              f"({last['t_sample']:.2f}s; the family loop is serial and exactly linear, SURVEY.md 8d; {ref.family_threads} omp thread(s) "
              f"for the per-node tasks, the fastest of 1/2/{ref.cores} here); ms_per_step is that EXTRAPOLATION, not a measured step")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * n_total / value, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": DATA,
        "config": {"workload": workload_label(name), "families_total": n_total, "W": ref.ranges[1] + 1,
                   "R": ref.ranges[3] - ref.ranges[2] + 1, "S": max(ref.ranges[1], ref.ranges[3]) + 1},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": ref.cores, "kind": ref.kind, "sample": sample,
                         "note": "unmodified reference compiled with OpenMP as its Makefile.in does; its lambda search is single-threaded "
                                 "over families (lambda.cpp:698-722), only the matrix build and the two child factors of a node are parallel"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# =====================================================================================================
# our arm
# =====================================================================================================
class GpuProblem:
    """One BASELINE configuration on this rank's slice of its table, through the C-ABI (cafe_b200.gpu = ctypes over libcafe_gpu.so)."""

    def __init__(self, name, rank, world, device, stream_ptr, comm_id=None):
        from cafe_b200 import gpu as cgpu
        from cafe_b200 import host as chost
        self.name, self.rank, self.world = name, rank, world
        self.cfg = bench_data.CONFIGS[name]
        self.newick = bench_data.config_tree(name)
        self.lam0 = bench_data.default_lambda(self.newick)
        counts, first0 = bench_data.config_slice(name, rank, world)
        self.counts = counts
        uniq, mult, first = bench_data.dedup(counts)
        self.n_unique = len(uniq)
        tree = chost.parse_tree(self.newick)
        rg = chost.init_family_size(self.cfg["max_size"])
        self.ranges = (rg["min"], rg["max"], rg["root_min"], rg["root_max"])
        self.R = self.ranges[3] - self.ranges[2] + 1
        self.prior = bench_data.root_prior(self.cfg["max_size"], self.ranges[2], self.R)  # the tables' own root distribution
        self.n = tree.n_nodes
        g = cgpu.CafeGpu(device)
        g.set_stream(stream_ptr)
        if world > 1:
            g.comm_init(comm_id, rank, world)  # ncclCommInitRank inside the library; K1 sharding + both exchanges from here on
        g.set_tree(tree.left, tree.right, tree.branchlength)
        g.set_ranges(*self.ranges)
        g.set_lnc_table(chost.lnc_table(max(self.ranges[1], self.ranges[3])))
        g.set_families(uniq, mult, first + first0)  # first_index = position in the whole table
        g.set_prior(self.prior)
        if name == "configs[3]":
            E = self.error_matrix(self.ranges[1] + 1)
            for k in range(tree.n_nodes // 2 + 1):
                g.set_error_model(k, E)
        self.g = g

    @staticmethod
    def error_matrix(dim):
        """errormodel with -1/0/+1 differences 0.05/0.9/0.05 (cafe/error_model.cpp:145-259), rows = observed, columns = true."""
        E = np.zeros((dim, dim))
        for t in range(dim):
            E[t, t] = 0.9
            if t > 0:
                E[t - 1, t] = 0.05
            if t + 1 < dim:
                E[t + 1, t] = 0.05
        E[0, 0] = 0.95
        E[dim - 1, dim - 1] = 0.95
        return E

    def rates(self, k):
        lam = lambda_schedule(self.lam0, k)
        mu = self.cfg["mu_ratio"] * lam if self.cfg["mu_ratio"] > 0 else -1.0
        if self.name == "configs[3]":  # `lambda -t` with four classes on branch subsets
            return lam * np.array([1.0, 1.3, 0.7, 1.1])[np.arange(self.n) % 4], np.full(self.n, mu)
        return np.full(self.n, lam), np.full(self.n, mu)


def measure(P, args, torch, dist, dev, stream, sampler=None):
    """Device-timed region (inputs resident) + end-to-end region (host rates in, host score out) of one configuration."""
    g, world = P.g, P.world
    out2 = torch.zeros(2, dtype=torch.float64, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for k in range(args.warmup):
        g.objective_device(*P.rates(k), out2.data_ptr())
    barrier()
    g.enable_timing(True)
    g.reset_launch_count()
    if sampler is not None:  # wait for the first nvidia-smi sample so that the timed region is covered
        t_wait = time.time()
        while sampler.mark() == 0 and time.time() - t_wait < 3.0:
            time.sleep(0.01)
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    barrier()
    mark0 = sampler.mark() if sampler else 0
    e0.record(stream)
    for k in range(args.steps):
        g.objective_device(*P.rates(args.warmup + k), out2.data_ptr())
    e1.record(stream)
    barrier()
    mark1 = (sampler.mark() + 1) if sampler else 0
    ms_total = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    launches = g.launch_count()
    k1_ms, x_ms, k2_ms, r_ms = g.timing_collect4()
    g.enable_timing(False)
    ms_per_step = float(ms_total.item()) / args.steps
    res = out2.cpu().numpy()
    last_score = float(res[0]) if np.isinf(res[1]) else float("-inf")

    # ---- end-to-end: the reference-facing call with HOST buffers in and out, every step ----
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        sc, fz = g.objective(*P.rates(args.warmup + k))  # host lambda/mu arrays in, host score out (sync inside)
    torch.cuda.synchronize()
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    families_total = P.cfg["families"]
    mean = lambda a: float(np.mean(a)) if len(a) else float("nan")  # noqa: E731
    return {
        "ms_per_step": ms_per_step, "value": families_total / (ms_per_step * 1e-3),
        "e2e_value": families_total / (float(t_e2e.item()) / args.steps),
        "k1_ms": mean(k1_ms), "exchange_ms": mean(x_ms), "k2_ms": mean(k2_ms), "reduce_ms": mean(r_ms),
        "launches": int(launches), "last_score": last_score, "e2e_score": float(sc),
        "flops_local": g.score_flops(), "marks": (mark0, mark1),
        "h2d": g.num_keys() * 48,  # BdKeyParams per distinct (int t, lambda, mu) key; rates are de-duplicated on the host
        "d2h": 16,
    }


def measure_pvalue_pass(P, torch, dist, n_samples=1000):
    """BASELINE configs[4]: the conditional distribution (n_samples Monte-Carlo families per root size, each pruned with its
    root size fixed: cafe/conditional_distribution.cpp:10-44) and the family-wide p-values of every family of the table
    (cafe/pvalue.cpp:143-154) through the C-ABI, host arrays out.  With N ranks the root sizes are split over the ranks and
    the rows all-gathered; every rank then does the p-values of its own families.  Wall clock, max over ranks."""
    from cafe_b200 import sharding
    g, world, rank = P.g, P.world, P.rank
    g.set_rates(*P.rates(0))
    g.build_matrices()
    g.synchronize()

    def wall(fn):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return out, float(t.item())

    def cond_dist():
        if world > 1:
            return sharding.conditional_distribution_sharded(g, n_samples, 7, rank, world)
        return g.conditional_distribution(n_samples, seed=7)

    cd, t_cd_first = wall(cond_dist)   # first call of the session: buffers, and with N > 1 the first all-gather of this size
    g.reset_launch_count()
    cd, t_cd = wall(cond_dist)         # steady state (the same seed: the same distribution)
    launches = g.launch_count()
    # wall clock of a host-driven pass (a dozen launches, allocations and frees per call) is exposed to host jitter: three steady-
    # state calls, the median is reported and all three are listed
    cd_runs = [t_cd] + [wall(cond_dist)[1] for _ in range(2)]
    t_cd = float(np.median(cd_runs))
    pv, t_pv_first = wall(lambda: g.pvalues(cd))   # first call of the session: allocates the root-row buffer (F x root rows doubles)
    g.reset_launch_count()
    pv, t_pv = wall(lambda: g.pvalues(cd))         # steady state (what a `report` after the first one pays)
    launches += g.launch_count()       # one distribution + one p-value pass
    pv_runs = [t_pv] + [wall(lambda: g.pvalues(cd))[1] for _ in range(2)]
    t_pv = float(np.median(pv_runs))
    per_family = g.score_flops() / max(1, P.n_unique)
    draws = P.R * n_samples
    return {"workload": f"conditional distribution: {n_samples} draws x {P.R} root sizes, then p-values of the "
                        f"{P.cfg['families']} families of the headline table (BASELINE configs[4])",
            "cd_s": t_cd, "cd_first_call_s": t_cd_first, "cd_s_runs": cd_runs, "pvalues_s": t_pv, "pvalues_first_call_s": t_pv_first,
            "pvalues_s_runs": pv_runs, "draws_per_s": draws / t_cd, "family_pvalues_per_s": P.cfg["families"] / t_pv,
            "cd_tflops_full_range_equivalent": draws * per_family / t_cd * 1e-12,
            "pvalues_tflops_full_range_equivalent": P.cfg["families"] * per_family / t_pv * 1e-12,
            "flops_note": "internal edges only, counted over the FULL range per family: the windowed kernel stops every K loop and "
                          "output pass at the largest window of a 96-family tile (as the reference stops at each family's own range), "
                          "so these figures are a work-equivalent rate and may exceed the DMMA peak; the kernel itself is the one "
                          "profiled in profiles/r2_k4_prune_fused2_windowed_ncu.txt",
            "gpu_launches": int(launches), "max_pvalue_mean": float(np.mean(pv)), "cd_checksum": float(cd.sum())}


def dgemm_peak(torch, dev):
    """fp64 roofline denominator: cuBLAS DGEMM on this GPU, this run (MEASURED_PEAKS.json carries no fp64 figure)."""
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
    return 2 * 6144 ** 3 / best * 1e-9


def traffic_of(name):
    tp = os.path.join(ROOT, "profiles", "k2_traffic.json")
    try:
        d = json.load(open(tp))
        return d.get(name, {}).get("dram_bytes_per_step") if name in d else (d.get("dram_bytes_per_step") if name == "configs[1]" else None)
    except Exception:
        return None


def roofline_of(m, peak_tflops, name, peaks):
    achieved = m["flops_local"] / (m["k2_ms"] * 1e-3) * 1e-12
    return {"bound": "tensor", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s", "frac": achieved / peak_tflops,
            "traffic": traffic_of(name), "kernel": "K2 batched pruning (DMMA.8x8x4 fp64), rank 0's shard",
            "flops_per_launch": m["flops_local"],
            "peak_source": "cuBLAS DGEMM 6144^3 fp64 measured in this run (MEASURED_PEAKS.json has no fp64 figure; tcgen05 has no "
                           "fp64 kind, the fp64 tensor pipe is DMMA)",
            "hbm_peak_gbs_measured": peaks.get("hbm_gbs")}


def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from cafe_b200 import gpu as cgpu

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    if bench_data.N_CHUNKS % world:
        raise SystemExit(f"bench.py: --gpus must divide {bench_data.N_CHUNKS}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    comm_id = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)  # barrier, max-over-ranks of the timings, hand-over of the NCCL id
        box = [cgpu.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        comm_id = box[0]
    # everything (our kernels, the library's NCCL collectives, the timing events) runs on one explicit stream
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    P = GpuProblem(HEADLINE, rank, world, local_rank, stream.cuda_stream, comm_id)
    m = measure(P, args, torch, dist, dev, stream, sampler if rank == 0 else None)
    clocks = sampler.stop(max(0, m["marks"][0] - 1), m["marks"][1]) if rank == 0 else None
    pv_pass = None
    if not args.no_sub:
        try:
            pv_pass = measure_pvalue_pass(P, torch, dist)
        except Exception as e:  # a secondary measurement must never take the headline down
            pv_pass = {"error": repr(e)}
    if rank != 0:
        P.g.close()
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = dgemm_peak(torch, dev)
    line = {
        "metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": DATA,
        "config": {
            "workload": workload_label(HEADLINE), "families_total": P.cfg["families"], "families_rank0": int(len(P.counts)),
            "unique_patterns_rank0": int(P.n_unique), "W": P.ranges[1] + 1, "R": P.R, "S": max(P.ranges[1], P.ranges[3]) + 1,
            "keys": P.g.num_keys(),
            "parallelism": f"families sharded x{world}" + (f", matrix build sharded x{world}, ncclAllGather of the matrices + "
                                                           f"ncclAllGather/ordered sum of the score inside libcafe_gpu.so" if world > 1 else ""),
            "l2": "no explicit flush: K2 streams ~1 GB of node-vector scratch per launch and the matrices (2 x 197 MB) are rewritten "
                  "every step, both larger than the 126 MB L2",
            "last_score": m["last_score"],
            "ms_breakdown_rank0": {"k1_matrix_build": m["k1_ms"], "matrix_allgather_and_transpose": m["exchange_ms"],
                                   "k2_pruning": m["k2_ms"], "score_reduction": m["reduce_ms"]},
        },
        "e2e": {"value": m["e2e_value"], "unit": UNIT, "h2d_bytes_per_step": m["h2d"], "d2h_bytes_per_step": m["d2h"],
                "what": "cafe_gpu_objective(host lambda/mu arrays) -> host score, per step; the family table is session state set "
                        "once, as after the reference's `load`"},
        "gpu_launches": m["launches"],
        "clocks": clocks,
        "roofline": roofline_of(m, peak, HEADLINE, peaks),
    }

    if world == 1 and not args.no_sub:
        subs = {}
        for name in ("configs[1]", "configs[3]"):
            if name == HEADLINE:
                continue
            try:
                Q = GpuProblem(name, 0, 1, local_rank, stream.cuda_stream)
                q = measure(Q, args, torch, dist, dev, stream)
                subs[name] = {"workload": bench_data.CONFIGS[name]["label"], "value": q["value"], "unit": UNIT,
                              "ms_per_step": q["ms_per_step"], "e2e": q["e2e_value"], "k1_ms": q["k1_ms"], "k2_ms": q["k2_ms"],
                              "W": Q.ranges[1] + 1, "R": Q.R, "keys": Q.g.num_keys(), "last_score": q["last_score"],
                              "roofline": roofline_of(q, peak, name, peaks)}
                Q.g.close()
            except Exception as e:  # a sub-configuration must never take the headline down
                subs[name] = {"error": repr(e)}
        line["configs"] = subs
    if pv_pass is not None:
        line.setdefault("configs", {})["configs[4]"] = pv_pass
        if world == 1 and "error" not in pv_pass and not args.no_cpu_baseline:
            try:
                ref4 = ReferenceCpu(HEADLINE)
                dps, dt = ref4.cond_dist_sample(P.R, 4)
                pv_pass["cpu_baseline"] = {"value": dps, "unit": "draws/s", "cores": ref4.cores, "kind": ref4.kind,
                                           "sample": f"the reference's pthreads cafe_conditional_distribution (one thread per block of root "
                                                     f"sizes, conditional_distribution.cpp:86-120) on all {P.R} root sizes x 4 draws instead of "
                                                     f"1000, {ref4.cores} threads ({dt:.1f}s), same tree, rates and ranges"}
                pv_pass["cd_speedup_vs_cpu"] = pv_pass["draws_per_s"] / dps
            except Exception as e:
                pv_pass["cpu_baseline"] = {"error": repr(e)}

    # ---- CPU baseline on this box's host cores (bounded sample of the same table) ----
    if world == 1 and not args.no_cpu_baseline:
        ref = ReferenceCpu(HEADLINE)
        ref.calibrate(P.counts[:8], P.lam0)
        r = ref.evaluate(P.counts[: args.cpu_sample], P.lam0, P.cfg["families"])
        line["cpu_baseline"] = {
            "value": r["value"], "unit": UNIT, "cores": ref.cores, "kind": ref.kind,
            "sample": f"{args.cpu_sample} of {P.cfg['families']} families (the first rows of the same table) + the full matrix build "
                      f"({r['t_matrices']:.2f}s matrices on {ref.cores} OpenMP threads, {r['t_sample']:.2f}s sample with "
                      f"{ref.family_threads} thread(s)), extrapolated linearly in F",
            "host_cpus": os.cpu_count()}
    print(json.dumps(line), flush=True)
    P.g.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=24)
    ap.add_argument("--ref-sample", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sub", action="store_true", help="skip the secondary configurations of the N=1 line")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank, local_rank, world = env_int("RANK", 0), env_int("LOCAL_RANK", 0), env_int("WORLD_SIZE", 1)
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
