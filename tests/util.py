"""Shared helpers for the parity tests: build a C-ABI context and the matching oracle inputs from the
same flat description."""
from __future__ import annotations

import numpy as np

import oracle
from cafe_b200 import gpu as cgpu
from cafe_b200 import host as chost

EXAMPLE_TREE = "(((chimp:6,human:6):81,(mouse:17,rat:17):70):6,dog:93)"


random_tree = oracle.random_tree
simulate_families = oracle.simulate_families


class Problem:
    """One likelihood problem described once, instantiated for the oracle and for the C-ABI."""

    def __init__(self, newick, counts, lam, mu=None, lambda_tree=None, ranges=None, prior_lambda=None,
                 err=None, mult=None, first=None):
        self.tree = chost.parse_tree(newick)
        self.otree = oracle.parse_newick(newick)
        t = self.tree
        self.counts = np.ascontiguousarray(counts, dtype=np.int32)
        mx = int(self.counts.max())
        if ranges is None:
            r = chost.init_family_size(mx)
            ranges = (r["min"], r["max"], r["root_min"], r["root_max"])
        self.ranges = ranges
        self.maxfs = max(ranges[1], ranges[3])
        n = t.n_nodes
        lam = np.atleast_1d(np.asarray(lam, dtype=np.float64))
        if lambda_tree is not None:
            _, ids = chost.parse_lambda_tree(newick, lambda_tree)
            ids = np.maximum(ids, 0)
        else:
            ids = np.zeros(n, dtype=np.int32)
        self.lam_node = lam[ids]
        if mu is None:
            self.mu_node = np.full(n, -1.0)
        else:
            mu = np.atleast_1d(np.asarray(mu, dtype=np.float64))
            self.mu_node = mu[ids]
        R = ranges[3] - ranges[2] + 1
        if prior_lambda is None:
            prior_lambda = max(1.0, float(self.counts[self.counts > 0].mean()) - 1.0)
        self.prior = oracle.prior_poisson(ranges[2], prior_lambda, 1000)[: max(R, 1)]
        self.err = err  # dict leaf-ordinal -> dense matrix, or None
        self.mult = mult
        self.first = first

    # ---- oracle side ----
    def oracle_mats(self):
        return oracle.node_matrices(self.otree, self.lam_node, self.mu_node, self.maxfs)

    def oracle_leaf_err(self):
        if not self.err:
            return None
        le = [None] * self.otree.n_nodes
        for leaf, M in self.err.items():
            le[2 * leaf] = M
        return le

    def oracle_score(self, want_L=True):
        return oracle.score(self.otree, self.oracle_mats(), self.counts, self.ranges, self.prior,
                            leaf_err=self.oracle_leaf_err(), want_L=want_L)

    # ---- C-ABI side ----
    def make_gpu(self, devices=None, comm=None, lo=None, hi=None):
        """devices=[...]: a multi-device leader context (cafe_gpu_create_multi); comm=(id, rank, world) + [lo, hi): one rank of
        a communicator holding that slice of the families."""
        g = cgpu.CafeGpu(devices=devices) if devices is not None else cgpu.CafeGpu(-1 if comm is None else comm[1])
        if comm is not None:
            g.comm_init(*comm)
        t = self.tree
        g.set_tree(t.left, t.right, t.branchlength)
        g.set_ranges(*self.ranges)
        g.set_lnc_table(chost.lnc_table(self.maxfs))
        if lo is None:
            g.set_families(self.counts, self.mult, self.first)
        else:
            first = np.arange(len(self.counts), dtype=np.int32) if self.first is None else self.first
            g.set_families(self.counts[lo:hi], None if self.mult is None else self.mult[lo:hi], first[lo:hi])
        g.set_prior(self.prior)
        if self.err:
            for leaf, M in self.err.items():
                g.set_error_model(leaf, M)
        g.set_rates(self.lam_node, self.mu_node)
        g.build_matrices()
        return g


def rel_err(a, b, floor=1e-300):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), floor)
