"""Family sharding across ranks: one process per GPU, torch.distributed for the plumbing.

Families are independent given the transition matrices (get_posterior has no cross-family state
except the running sum and the first-zero exception, cafe/lambda.cpp:698-722), so every rank holds
all D matrices and a contiguous slice of the unique patterns.  The only exchange step of one
objective evaluation is the reduction of {partial score, first zero-likelihood family index}:
ONE collective on a 2-double device buffer (sum of the first, min of the second).
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n_items: int, world_size: int, rank: int):
    """Contiguous, balanced [lo, hi) slice of n_items for `rank` (sizes differ by at most one)."""
    base, extra = divmod(n_items, world_size)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def shard_families(counts: np.ndarray, multiplicity, first_index, world_size: int, rank: int):
    lo, hi = shard_bounds(len(counts), world_size, rank)
    m = None if multiplicity is None else multiplicity[lo:hi]
    f = np.arange(lo, hi, dtype=np.int32) if first_index is None else first_index[lo:hi]
    return counts[lo:hi], m, f


def reduce_score(local2, group=None):
    """local2: torch tensor [2] = (partial score, min first-index of a zero family or +inf) on this
    rank's device (CPU tensors with gloo work too).  Returns (score, first_zero) identical on every
    rank, using a single all_gather of 2 doubles per rank."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        s, z = local2[0], local2[1]
    else:
        ws = dist.get_world_size(group)
        gathered = torch.empty(ws * 2, dtype=local2.dtype, device=local2.device)
        dist.all_gather_into_tensor(gathered, local2.contiguous(), group=group)
        g = gathered.view(ws, 2)
        s, z = g[:, 0].sum(), g[:, 1].min()
    return s, z


class _DeviceBuffer:
    """A device allocation owned by the C-ABI library, exposed to torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr: int, n_doubles: int):
        self.__cuda_array_interface__ = {"shape": (n_doubles,), "typestr": "<f8", "data": (ptr, False), "version": 2}


def gather_chunks(full, rank: int, world_size: int, group=None):
    """In-place all-gather over a flat tensor of world_size equal chunks: chunk `rank` holds this rank's data on entry,
    every chunk is filled on return.  (CPU tensors with gloo work too.)"""
    import torch.distributed as dist

    chunk = full.numel() // world_size
    dist.all_gather_into_tensor(full, full[rank * chunk:(rank + 1) * chunk], group=group)
    return full


def objective_sharded(g, lam_node, mu_node, out2, rank: int, world_size: int, device, group=None):
    """One objective evaluation with BOTH kernels sharded: every rank builds ceil(D/world) of the D distinct transition
    matrices (K1), two in-place NCCL all-gathers over NVLink (M and its transposed copy) give every rank all of them, then
    K2+K3 run on this rank's families and the 2-double reduction follows (reduce_score).  The matrices of BASELINE configs[1]
    are 20 x 0.5 MB x 2, those of configs[2] 98 x 2 MB x 2.  g must have had set_key_shard(rank, world_size)."""
    import torch

    g.set_rates(lam_node, mu_node)
    g.build_matrices()
    pm, pt, dpk, kpr = g.matrix_storage()
    n = dpk * kpr * world_size
    for ptr in (pm, pt):
        gather_chunks(torch.as_tensor(_DeviceBuffer(ptr, n), device=device), rank, world_size, group)
    g.matrices_exchanged()
    g.score_device(out2.data_ptr())
    return reduce_score(out2, group)


def finish_score(s, z):
    """Host-side decode of reduce_score's result: (-inf, index) when some family had zero likelihood."""
    s, z = float(s), float(z)
    if np.isinf(z):
        return s, -1
    return -np.inf, int(z)
