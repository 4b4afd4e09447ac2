"""Synthetic, in-model gene-family tables for the benchmark configurations (BASELINE.json `configs`).

Trees: random ultrametric binary trees with INTEGER branch lengths (coalescent-style merging, gaps
uniform in {1,2,3}) — SURVEY.md §8d.  Families: simulated from the birth–death model itself at
lambda0 = 0.25/depth, walking down the tree and drawing each child size from the parent's row of the
branch's transition matrix.  The matrices come from the product's own K1 kernel through the C-ABI
(cafe_gpu_get_matrix); the draws are numpy on the host.  Nothing here touches oracle/.
"""
from __future__ import annotations

import numpy as np

from . import gpu as cgpu
from . import host as chost


def random_tree(n_leaves: int, seed: int = 1, max_gap: int = 3) -> str:
    rng = np.random.RandomState(seed)
    nodes = [(f"s{i}", 0) for i in range(n_leaves)]
    h = 0
    while len(nodes) > 1:
        h += int(rng.randint(1, max_gap + 1))
        i, j = sorted(rng.choice(len(nodes), 2, replace=False))
        a, b = nodes[i], nodes[j]
        new = (f"({a[0]}:{h - a[1]},{b[0]}:{h - b[1]})", h)
        nodes = [x for k, x in enumerate(nodes) if k not in (i, j)] + [new]
    return nodes[0][0]


def tree_depth(tree) -> float:
    v, d = 0, 0.0
    while tree.parent[v] >= 0:
        d += tree.branchlength[v]
        v = tree.parent[v]
    return d


def prefix_order(tree):
    order, st = [], [tree.root]
    while st:
        v = st.pop()
        order.append(v)
        if tree.left[v] >= 0:
            st.append(tree.right[v])
            st.append(tree.left[v])
    return order


def simulate_table(newick: str, n_families: int, max_size: int, lam0: float | None = None, mu0: float | None = None,
                   seed: int = 10, device: int = -1):
    """Return (counts[F][n_leaves] int32 in leaf order, lam0) with observed max == max_size exactly."""
    tree = chost.parse_tree(newick)
    if lam0 is None:
        lam0 = 0.25 / tree_depth(tree)
    rg = chost.init_family_size(max_size)
    S = max(rg["max"], rg["root_max"]) + 1
    g = cgpu.CafeGpu(device)
    try:
        g.set_tree(tree.left, tree.right, tree.branchlength)
        g.set_ranges(rg["min"], rg["max"], rg["root_min"], rg["root_max"])
        g.set_lnc_table(chost.lnc_table(S - 1))
        n = tree.n_nodes
        g.set_rates(np.full(n, lam0), np.full(n, -1.0 if mu0 is None else mu0))
        g.build_matrices()
        cdf = {}
        for v in range(n):
            if v == tree.root:
                continue
            key = int(tree.branchlength[v])
            if key not in cdf:
                cdf[key] = np.cumsum(g.get_matrix(v), axis=1)
    finally:
        g.close()
    rng = np.random.RandomState(seed)
    order = prefix_order(tree)
    kept = []
    total = 0
    have_max = False
    while total < n_families or not have_max:
        B = max(4096, int(n_families * 0.4))
        # root sizes: mostly small families plus a flat tail that reaches max_size
        small = 1 + rng.poisson(8.0, size=B)
        tail = rng.randint(1, max_size + 1, size=B)
        root = np.where(rng.random_sample(B) < 0.85, small, tail)
        sizes = np.zeros((B, n), dtype=np.int64)
        sizes[:, tree.root] = np.minimum(root, S - 1)
        for v in order:
            if v == tree.root:
                continue
            c = cdf[int(tree.branchlength[v])]
            par = sizes[:, tree.parent[v]]
            u = rng.random_sample(B)
            child = np.empty(B, dtype=np.int64)
            for p in np.unique(par):
                idx = np.where(par == p)[0]
                child[idx] = np.searchsorted(c[p], u[idx], side="left")
            sizes[:, v] = np.minimum(child, S - 1)
        leaves = sizes[:, 0::2]
        ok = leaves.max(axis=1) <= max_size
        leaves = leaves[ok]
        hit = leaves.max(axis=1) == max_size
        if hit.any() and not have_max:
            first = np.where(hit)[0][0]
            kept.insert(0, leaves[first:first + 1])  # make sure the table's max is exactly max_size
            total += 1
            have_max = True
        kept.append(leaves)
        total += len(leaves)
    counts = np.concatenate(kept, axis=0)[:n_families].astype(np.int32)
    assert counts.max() == max_size
    return counts, lam0


def dedup(counts: np.ndarray):
    """Hash-style duplicate detection with the reference's `ref` semantics (first occurrence wins,
    cafe/cafe_family.c:9-34): returns (unique_counts in first-occurrence order, multiplicity, first_index)."""
    _, first, inverse, mult = np.unique(counts, axis=0, return_index=True, return_inverse=True, return_counts=True)
    order = np.argsort(first)
    return counts[first[order]], mult[order].astype(np.int32), first[order].astype(np.int32)
