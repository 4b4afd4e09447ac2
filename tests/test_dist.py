"""CPU, world_size 2, gloo: the multi-rank plumbing of one objective evaluation — contiguous family
shards, one collective on {partial score, first zero family}, identical result on every rank."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from cafe_b200 import sharding

EX_TREE = "(((chimp:6,human:6):81,(mouse:17,rat:17):70):6,dog:93)"


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 100, 50000, 200001):
        for ws in (1, 2, 3, 8):
            cuts = [sharding.shard_bounds(n, ws, r) for r in range(ws)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(ws - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1


def _problem(lam):
    t = oracle.parse_newick(EX_TREE)
    rng = np.random.RandomState(4)
    base = rng.randint(1, 20, size=(41, 1))
    counts = np.maximum(0, base + rng.randint(-2, 3, size=(41, 5))).astype(np.int32)
    ranges = (0, 72, 1, 30)
    mats = oracle.node_matrices(t, [lam] * t.n_nodes, [-1.0] * t.n_nodes, 72)
    prior = oracle.prior_poisson(1, 8.0, 1000)[:30]
    return t, counts, ranges, mats, prior


def _worker(rank, world, port, lam, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    t, counts, ranges, mats, prior = _problem(lam)
    first = np.arange(len(counts), dtype=np.int32) + 1000
    c, m, f = sharding.shard_families(counts, None, first, world, rank)
    # the per-rank partial a GPU context would leave in its 2-double device buffer (cafe_gpu_objective_device)
    o = oracle.score(t, mats, c, ranges, prior)
    zero = o["maxlik"] == 0
    partial = float(o["logpost"][~zero].sum())
    zidx = float(f[zero].min()) if zero.any() else float("inf")
    s, z = sharding.reduce_score(torch.tensor([partial, zidx], dtype=torch.float64))
    out[rank] = sharding.finish_score(s, z)
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("lam,expect_zero", [(0.005, False), (0.011, True)])
def test_two_rank_reduction_matches_single_rank(lam, expect_zero):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), lam, out), nprocs=world, join=True)
    t, counts, ranges, mats, prior = _problem(lam)
    full = oracle.score(t, mats, counts, ranges, prior)
    assert out[0] == out[1]
    score, fz = out[0]
    if expect_zero:
        assert score == -np.inf and fz == 1000 + full["first_zero"]
    else:
        assert fz == -1 and abs(score - full["score"]) < 1e-9


def test_reduce_score_without_process_group():
    s, z = sharding.reduce_score(torch.tensor([-12.5, float("inf")], dtype=torch.float64))
    assert sharding.finish_score(s, z) == (-12.5, -1)
    s, z = sharding.reduce_score(torch.tensor([-12.5, 7.0], dtype=torch.float64))
    assert sharding.finish_score(s, z) == (-np.inf, 7)


def _gather_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # the matrix buffers of a sharded K1: D = 5 keys -> keys_per_rank = 3, buffers hold 3 * 2 matrices of 4 doubles;
    # every rank has filled its own chunk only (key k holds the value 10 + k), the rest is stale
    keys_per_rank, per_key, D = 3, 4, 5
    full = torch.full((keys_per_rank * world * per_key,), -1.0, dtype=torch.float64)
    for k in range(rank * keys_per_rank, min(D, (rank + 1) * keys_per_rank)):
        full[k * per_key:(k + 1) * per_key] = 10.0 + k
    sharding.gather_chunks(full, rank, world)
    out[rank] = full.view(-1, per_key)[:D, 0].tolist()
    dist.destroy_process_group()


def test_in_place_chunk_gather_gives_every_rank_all_keys():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_gather_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert out[0] == out[1] == [10.0, 11.0, 12.0, 13.0, 14.0]


def _rows_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_rows, n = 7, 3   # 7 root sizes over 2 ranks: shards of 4 and 3 rows
    full = np.arange(n_rows * n, dtype=np.float64).reshape(n_rows, n)
    lo, hi = sharding.shard_bounds(n_rows, world, rank)
    out[rank] = sharding.gather_rows(full[lo:hi], n_rows, rank, world).tolist()
    dist.destroy_process_group()


def test_row_sharded_matrix_gather():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_rows_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    expect = np.arange(21, dtype=np.float64).reshape(7, 3).tolist()
    assert out[0] == expect and out[1] == expect


# ------------------------------------------------------------------ likelihood-ratio test, families sharded
LRT_TREE = "(((chimp:6.6,human:6.6):81.2,(mouse:17.4,rat:17.4):70.4):6.9,dog:93.7)"   # fractional lengths: the first-family quirk matters


class _OracleLrtContext:
    """Stands in for a C-ABI context that holds one shard: likelihood_ratio_test with the ABI's semantics (tested 0 / 1 / 2),
    computed by the oracle (which tests/test_oracle.py pins against the compiled reference)."""

    def __init__(self, tree, mats, lam, mu, counts, ranges):
        self.t, self.mats, self.lam, self.mu, self.counts, self.ranges = tree, mats, lam, mu, counts, ranges

    def likelihood_ratio_test(self, tested=None, lengthened_mu=None):
        t = self.t
        F = len(self.counts)
        tested = np.ones(F, dtype=np.uint8) if tested is None else tested
        base = np.zeros(F); best = np.zeros((t.n_nodes, F)); steps = np.zeros((t.n_nodes, F), dtype=np.int32)
        bl = np.array(t.branchlength, dtype=np.float64)
        first_done = False
        for f in range(F):
            base[f] = oracle.prune(t, self.mats, self.counts[f], self.ranges).max()
            if not tested[f]:
                best[:, f] = base[f]; best[t.root, f] = -1
                continue
            if not first_done and tested[f] == 2:
                bl = np.floor(bl)             # not the table's first tested family: truncated lengths from the start
            first_done = True
            _, b, s = oracle.lrt_family(t, self.mats, self.lam, self.mu if lengthened_mu is None else lengthened_mu, bl,
                                        self.counts[f], self.ranges)
            best[:, f] = b; steps[:, f] = s
        return base, best, steps


def _lrt_problem():
    t = oracle.parse_newick(LRT_TREE)
    rng = np.random.RandomState(9)
    counts = np.maximum(0, rng.randint(1, 14, size=(11, 1)) + rng.randint(-3, 4, size=(11, 5))).astype(np.int32)
    counts[6, 1] += 20
    ranges = (0, 60, 1, 40)
    lam = np.full(t.n_nodes, 0.005); mu = np.full(t.n_nodes, -1.0)
    mats = oracle.node_matrices(t, lam, mu, 60)
    tested = np.ones(11, dtype=np.uint8)
    tested[[0, 1, 7]] = 0                      # the first tested family (2) sits in shard 0, shard 1 starts with a tested one
    return t, counts, ranges, lam, mu, mats, tested


def _lrt_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    t, counts, ranges, lam, mu, mats, tested = _lrt_problem()
    lo, hi = sharding.shard_bounds(len(counts), world, rank)
    g = _OracleLrtContext(t, mats, lam, mu, counts[lo:hi], ranges)
    out[rank] = sharding.likelihood_ratio_test_sharded(g, tested[lo:hi], len(counts), rank, world)
    dist.destroy_process_group()


def test_likelihood_ratio_test_sharded_equals_single_rank():
    t, counts, ranges, lam, mu, mats, tested = _lrt_problem()
    single = _OracleLrtContext(t, mats, lam, mu, counts, ranges).likelihood_ratio_test(tested)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_lrt_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    for r in range(2):
        for a, b in zip(out[r], single):
            assert np.array_equal(a, b)
    assert single[2].max() >= 2
    # without the marker the second shard's first family would start from the parsed lengths: make sure the case is exercised
    lo, hi = sharding.shard_bounds(len(counts), 2, 1)
    wrong = _OracleLrtContext(t, mats, lam, mu, counts[lo:hi], ranges).likelihood_ratio_test(tested[lo:hi])
    assert not np.array_equal(wrong[1], single[1][:, lo:hi])
