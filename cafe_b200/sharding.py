"""Family sharding across ranks: one process per GPU, torch.distributed for the plumbing.

Families are independent given the transition matrices (get_posterior has no cross-family state
except the running sum and the first-zero exception, cafe/lambda.cpp:698-722), so every rank holds
all D matrices and a contiguous slice of the unique patterns.  The exchange steps of one objective evaluation
(all-gather of the K1-sharded matrices, reduction of {partial score, first zero-likelihood family index}) run inside
libcafe_gpu.so over NCCL (csrc/comm.cu; attach_comm below hands over the communicator id).  The torch.distributed helpers
here cover what sits around it: slicing the table, the host-side form of the reduction (gloo, for CPU tests), and the
row/family gathers of the conditional distribution and the likelihood-ratio test.
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


def attach_comm(g, rank: int, world_size: int, group=None):
    """Give context `g` an NCCL communicator INSIDE the C-ABI library (cafe_gpu_comm_init): rank 0 draws the unique id
    (cafe_gpu_comm_unique_id), torch.distributed — any backend — only carries its 128 bytes to the other ranks.  From then on
    g.objective / g.score / g.objective_device shard K1 over the ranks, all-gather the matrices, reduce the score and return the
    same result on every rank; every collective runs on the context's own stream, ordered with its kernels."""
    import torch.distributed as dist

    from . import gpu as cgpu

    if world_size == 1:
        return
    box = [cgpu.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    g.comm_init(box[0], rank, world_size)


def gather_rows(local_rows, n_rows: int, rank: int, world_size: int, group=None):
    """All-gather of a row-sharded matrix: rank r holds rows shard_bounds(n_rows, world, r) as a [rows_r][n] array; every rank
    gets the full [n_rows][n] array.  Shards are padded to the largest one, so one all_gather_into_tensor suffices."""
    import torch
    import torch.distributed as dist

    local = torch.as_tensor(np.ascontiguousarray(local_rows, dtype=np.float64))
    n = local.shape[1]
    if not (dist.is_available() and dist.is_initialized()) or world_size == 1:
        return local.numpy()
    per = (n_rows + world_size - 1) // world_size
    use_cuda = dist.get_backend(group) == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if use_cuda else torch.device("cpu")
    mine = torch.zeros((per, n), dtype=torch.float64, device=dev)
    mine[: local.shape[0]] = local.to(dev)
    full = torch.empty((world_size * per, n), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(full, mine, group=group)
    full = full.cpu().numpy().reshape(world_size, per, n)
    return np.concatenate([full[r, : shard_bounds(n_rows, world_size, r)[1] - shard_bounds(n_rows, world_size, r)[0]] for r in range(world_size)])


def conditional_distribution_sharded(g, n_samples: int, seed: int, rank: int, world_size: int, group=None):
    """The conditional distribution (K4) with its R root-size rows split over the ranks — every rank simulates and prunes
    R/world x n_samples families, one all-gather of the sorted rows (R x n_samples doubles, 4 MB at BASELINE configs[4]) gives
    every rank the whole matrix for the p-values of its own families (cafe_gpu_pvalues).  The device RNG is keyed by
    (seed, root size, trial, node): the result does not depend on the number of ranks."""
    lo, hi = shard_bounds(g.R, world_size, rank)
    return gather_rows(g.conditional_distribution_rows(n_samples, lo, hi, seed=seed), g.R, rank, world_size, group)


def likelihood_ratio_test_sharded(g, tested_local, n_total: int, rank: int, world_size: int, lengthened_mu=None, group=None):
    """The branch-stretch likelihood-ratio test (cafe_gpu_likelihood_ratio_test) with the families split over the ranks like the
    score: every rank tests the families of its shard (the context `g` holds exactly those), no collective on the data path; one
    all-gather of a flag per rank settles which shard owns the table's first tested family - only that family starts from the
    parsed branch lengths (cafe/cafe_main.c:350,390), so the other shards mark theirs with 2 - and one all-gather brings the rows
    together.  tested_local: uint8 per family of this shard (1 test, 0 skip).  Returns (base [n_total], best [n_nodes][n_total],
    steps [n_nodes][n_total]) on every rank."""
    import torch
    import torch.distributed as dist

    tested_local = np.ascontiguousarray(tested_local, dtype=np.uint8).copy()
    multi = dist.is_available() and dist.is_initialized() and world_size > 1
    if multi:
        use_cuda = dist.get_backend(group) == "nccl"
        dev = torch.device("cuda", torch.cuda.current_device()) if use_cuda else torch.device("cpu")
        mine = torch.tensor([1 if tested_local.any() else 0], dtype=torch.int32, device=dev)
        flags = torch.empty(world_size, dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(flags, mine, group=group)
        if bool(flags[:rank].any().item()):
            tested_local[tested_local == 1] = 2          # an earlier shard owns the first tested family
    base, best, steps = g.likelihood_ratio_test(tested_local, lengthened_mu)
    if not multi:
        return base, best, steps
    # rows = families for the gather: [F_local][1 + 2 * n_nodes]
    local = np.concatenate([base[:, None], best.T, steps.T.astype(np.float64)], axis=1)
    full = gather_rows(local, n_total, rank, world_size, group)
    n_nodes = best.shape[0]
    return full[:, 0].copy(), full[:, 1:1 + n_nodes].T.copy(), full[:, 1 + n_nodes:].T.astype(np.int32)


def finish_score(s, z):
    """Host-side decode of reduce_score's result: (-inf, index) when some family had zero likelihood."""
    s, z = float(s), float(z)
    if np.isinf(z):
        return s, -1
    return -np.inf, int(z)
