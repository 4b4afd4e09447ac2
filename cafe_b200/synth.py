"""Synthetic, in-model gene-family tables for the benchmark configurations (BASELINE.json `configs`).

The generator itself lives in bench_data.py at the repository root (pure numpy, shared by both arms of bench.py so that
they time the same families); this module is the package-side door to it for tests and tools.
"""
from __future__ import annotations

import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from bench_data import dedup, random_tree, simulate_table  # noqa: E402,F401
from bench_data import tree_depth as _newick_depth  # noqa: E402


def tree_depth(tree) -> float:
    """Root-to-leaf depth of a parsed host tree (cafe_b200.host.parse_tree) or of a newick string."""
    if isinstance(tree, str):
        return _newick_depth(tree)
    v, d = 0, 0.0
    while tree.parent[v] >= 0:
        d += tree.branchlength[v]
        v = tree.parent[v]
    return d
