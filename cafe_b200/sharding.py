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


def finish_score(s, z):
    """Host-side decode of reduce_score's result: (-inf, index) when some family had zero likelihood."""
    s, z = float(s), float(z)
    if np.isinf(z):
        return s, -1
    return -np.inf, int(z)
