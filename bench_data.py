"""Synthetic, in-model gene-family tables for the BASELINE.json configurations — pure numpy, no GPU, no oracle.

Both arms of bench.py (ours and `--impl reference`), the full-size tests and the tools draw their tables from here, so they
time and check the SAME families.  Nothing in this file belongs to the product path or to the oracle.

Trees: random ultrametric binary trees with INTEGER branch lengths (coalescent-style merging, gaps uniform in {1,2,3}) —
SURVEY.md §8d.  Families: simulated from the linear birth–death process itself, walking down the tree.  For one lineage and
time t the number of descendants is 0 with probability alpha and otherwise geometric with parameter beta
(alpha, beta as in libtree/birthdeath.c:246-262), so the child size given parent size s is exactly
    k ~ Binomial(s, 1 - alpha) surviving lineages,   child = k + NegBinomial(k, 1 - beta)
— the same distribution as row s of the transition matrix, without building the matrix.
"""
from __future__ import annotations

import numpy as np


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


class _Tree:
    """children / branch lengths in parse order; `leaves` in left-to-right order = the product's leaf order (node 2k)."""

    def __init__(self, newick: str):
        s = newick.strip().rstrip(";")
        self.kids, self.t, self.leaves = [], [], []
        pos = 0

        def node():
            nonlocal pos
            me = len(self.kids)
            self.kids.append([])
            self.t.append(0.0)
            if s[pos] == "(":
                pos += 1
                while True:
                    self.kids[me].append(node())
                    if s[pos] == ",":
                        pos += 1
                        continue
                    pos += 1  # ')'
                    break
            else:
                self.leaves.append(me)
            while pos < len(s) and s[pos] not in ":,)":
                pos += 1
            if pos < len(s) and s[pos] == ":":
                pos += 1
                st = pos
                while pos < len(s) and s[pos] not in ",)":
                    pos += 1
                self.t[me] = float(s[st:pos])
            return me

        self.root = node()

    def depth(self) -> float:
        v, d = self.root, 0.0
        while self.kids[v]:
            v = self.kids[v][0]
            d += self.t[v]
        return d


def tree_depth(newick: str) -> float:
    return _Tree(newick).depth()


def default_lambda(newick: str) -> float:
    """lambda0 = 0.25 / depth: sizes stay well inside the range on the way down, likelihoods never underflow."""
    return 0.25 / _Tree(newick).depth()


def _alpha_beta(lam: float, mu: float, t: float):
    if mu < 0 or mu == lam:
        a = lam * t / (1 + lam * t)
        return a, a
    e = np.exp((lam - mu) * t)
    return mu * (e - 1) / (lam * e - mu), lam * (e - 1) / (lam * e - mu)


def simulate_table(newick: str, n_families: int, max_size: int, lam0: float | None = None, mu0: float | None = None, seed: int = 10):
    """(counts[F][n_leaves] int32 in leaf order, lam0): F families whose largest observed size is exactly max_size."""
    tr = _Tree(newick)
    if lam0 is None:
        lam0 = 0.25 / tr.depth()
    mu = -1.0 if mu0 is None else mu0
    rng = np.random.RandomState(seed)
    kept, total, have_max = [], 0, False
    order = []
    st = [tr.root]
    while st:
        v = st.pop()
        order.append(v)
        st.extend(reversed(tr.kids[v]))
    parent = {c: v for v in range(len(tr.kids)) for c in tr.kids[v]}
    best_short = None
    for attempt in range(200):
        if total >= n_families and have_max:
            break
        B = max(4096, int(n_families * 0.4))
        # root sizes: mostly small families plus a flat tail that reaches max_size
        small = 1 + rng.poisson(SMALL_MEAN, size=B)
        tail = rng.randint(1, max_size + 1, size=B)
        root = np.where(rng.random_sample(B) < SMALL_FRACTION, small, tail)
        sizes = np.zeros((B, len(tr.kids)), dtype=np.int64)
        sizes[:, tr.root] = root
        for v in order:
            if v == tr.root:
                continue
            a, b = _alpha_beta(lam0, mu, float(int(tr.t[v])))  # the (int) branch length of the matrix key, cafe_tree.c:376
            par = sizes[:, parent[v]]
            k = rng.binomial(par, 1.0 - a)
            extra = np.zeros(B, dtype=np.int64)
            pos = k > 0
            extra[pos] = rng.negative_binomial(k[pos], 1.0 - b)
            sizes[:, v] = k + extra
        leaves = sizes[:, tr.leaves]
        mx = leaves.max(axis=1)
        ok = mx <= max_size
        leaves, mx = leaves[ok], mx[ok]
        hit = mx == max_size
        if hit.any() and not have_max:
            first = int(np.where(hit)[0][0])
            kept.insert(0, leaves[first:first + 1])  # make sure the table's max is exactly max_size
            total += 1
            have_max = True
        elif not have_max and len(mx) and (best_short is None or mx.max() > best_short.max()):
            best_short = leaves[int(np.argmax(mx))].copy()
        if total < n_families:
            kept.append(leaves)
            total += len(leaves)
        if attempt >= 8 and not have_max and best_short is not None:
            # no family reached max_size on its own: stretch the largest one (one slightly out-of-model family)
            best_short[int(np.argmax(best_short))] = max_size
            kept.insert(0, best_short[None, :])
            total += 1
            have_max = True
    counts = np.concatenate(kept, axis=0)[:n_families].astype(np.int32)
    assert counts.max() == max_size and len(counts) == n_families
    return counts, lam0


SMALL_FRACTION, SMALL_MEAN = 0.85, 8.0  # root sizes: 85 % 1 + Poisson(8), 15 % uniform on 1..max_size


def root_prior(max_size: int, root_min: int, n: int) -> np.ndarray:
    """The root-size distribution the tables are drawn from, prior[i] for root size root_min + i — what the reference's
    `rootdist` / cafe_set_prior_rfsize_empirical stand for.  (A pure Poisson prior underflows to 0 long before size 400:
    every large family would then score log 0.)"""
    from scipy.special import gammaln
    sz = root_min + np.arange(n, dtype=np.float64)
    k = sz - 1.0
    pois = np.where(k >= 0, np.exp(k * np.log(SMALL_MEAN) - SMALL_MEAN - gammaln(np.maximum(k, 0) + 1.0)), 0.0)
    flat = np.where((sz >= 1) & (sz <= max_size), 1.0 / max_size, 0.0)
    return SMALL_FRACTION * pois + (1.0 - SMALL_FRACTION) * flat


def dedup(counts: np.ndarray):
    """Duplicate detection with the reference's `ref` semantics (first occurrence wins, cafe/cafe_family.c:9-34):
    returns (unique_counts in first-occurrence order, multiplicity, first_index)."""
    _, first, inverse, mult = np.unique(counts, axis=0, return_index=True, return_inverse=True, return_counts=True)
    order = np.argsort(first)
    return counts[first[order]], mult[order].astype(np.int32), first[order].astype(np.int32)


# ------------------------------------------------------------------------------------------------ BASELINE configurations
# name -> (taxa, families, max size, mu / lambda or 0, tree seed, table seed).  Tables are built in N_CHUNKS equal chunks with
# their own seeds, so that rank r of `world` (world divides N_CHUNKS) can build exactly its slice of the SAME table.
N_CHUNKS = 8
CONFIGS = {
    "configs[1]": dict(taxa=20, families=50000, max_size=200, mu_ratio=0.0, tree_seed=1, seed=10,
                       label="50000 families x 20 taxa, max size 200, single lambda (BASELINE configs[1])"),
    "configs[2]": dict(taxa=50, families=200000, max_size=400, mu_ratio=0.8, tree_seed=1, seed=20,
                       label="200000 families x 50 taxa, max size 400, lambda and mu (BASELINE configs[2])"),
    "configs[3]": dict(taxa=100, families=100000, max_size=200, mu_ratio=0.0, tree_seed=1, seed=30,
                       label="100000 families x 100 taxa, max size 200, 4 lambda classes, error model on every leaf (BASELINE configs[3])"),
}


def config_tree(name: str) -> str:
    c = CONFIGS[name]
    return random_tree(c["taxa"], c["tree_seed"])


def config_chunk(name: str, chunk: int) -> np.ndarray:
    """Chunk `chunk` (0..N_CHUNKS-1) of the configuration's table; the chunks concatenated are the table."""
    c = CONFIGS[name]
    nw = config_tree(name)
    lam0 = default_lambda(nw)
    counts, _ = simulate_table(nw, c["families"] // N_CHUNKS, c["max_size"], lam0, c["mu_ratio"] * lam0 if c["mu_ratio"] > 0 else None,
                               seed=c["seed"] + chunk)
    return counts


def config_slice(name: str, rank: int, world: int):
    """(counts of rank `rank`'s contiguous slice, index of its first family in the whole table)."""
    assert N_CHUNKS % world == 0, "world size must divide the number of table chunks"
    per = N_CHUNKS // world
    parts = [config_chunk(name, c) for c in range(rank * per, (rank + 1) * per)]
    return np.concatenate(parts, axis=0), rank * per * (CONFIGS[name]["families"] // N_CHUNKS)
