"""Error budget of a split-precision (Ozaki) tensor-core route for K2 — CPU study, numpy only.

The north star names "tcgen05 ... fp32/split where tolerance allows".  tcgen05.mma has no fp64 kind, so fp64 results on the
5th-generation tensor cores mean slicing both operands into short integers (int8 `kind::i8`, exact int32 accumulation) and
recombining the slice products in fp64 (Ozaki scheme): with S slices of BITS bits per operand and the slice pairs
s_a + s_b < S kept, one fp64 GEMM costs S (S + 1) / 2 int8 GEMMs.

What decides S here is not the 53-bit mantissa but the DYNAMIC RANGE inside one row: a family's node vector spans hundreds of
orders of magnitude over the sizes, slices are taken relative to the row's largest entry, and everything more than
S * BITS bits below it is dropped.  This script prunes real families of BASELINE configs[1] (the bench table) with every
internal-edge GEMM emulated that way — per-family-row and per-matrix-row power-of-two scaling, exact integer products,
fp64 recombination — and reports the relative error of the root likelihood vector's maximum (what the score uses) against
plain fp64, for a range of S.  Output: one JSON line per S (profiles/r2_ozaki_error_study.jsonl).
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_data  # noqa: E402
import oracle  # noqa: E402  (CPU matrices; this is a study tool, not the product path)

BITS = 7          # magnitude bits of a signed int8 slice
N_FAM = int(os.environ.get("OZAKI_FAMILIES", 768))


def slice_rows(X, S):
    """X >= 0, [rows][K].  Returns (slices [S][rows][K] int64 with values < 2^BITS, scale [rows]) with
    X ~= sum_s slices[s] * scale * 2^(-BITS (s+1)), truncating below 2^(-BITS S) of each row's largest entry."""
    mx = X.max(axis=1)
    e = np.where(mx > 0, np.ceil(np.log2(np.where(mx > 0, mx, 1.0))), 0.0)     # row max < 2^e
    scale = np.ldexp(1.0, e.astype(np.int64))
    rem = X / scale[:, None]                                                    # in [0, 1)
    out = []
    for _ in range(S):
        rem = rem * (1 << BITS)
        q = np.floor(rem)
        out.append(q.astype(np.int64))
        rem = rem - q
    return np.array(out), scale


def ozaki_gemm(A, B, S):
    """C[f][i] = sum_j A[f][j] B[i][j] with both operands sliced (A by family row, B by matrix row)."""
    As, sa = slice_rows(A, S)
    Bs, sb = slice_rows(B, S)
    C = np.zeros((A.shape[0], B.shape[0]))
    for p in range(S):
        for q in range(S - p):
            C += (As[p] @ Bs[q].T).astype(np.float64) * 2.0 ** (-BITS * (p + q + 2))   # exact int32-range products
    return C * sa[:, None] * sb[None, :]


def prune(tree, mats, counts, W, R, root_min, gemm):
    """Felsenstein pruning over all families at once (cafe_tree.c:191-271); gemm(A, B) -> A @ B.T."""
    F = counts.shape[0]

    def vec(v):
        if tree.left[v] < 0:
            L = np.zeros((F, W))
            L[np.arange(F), counts[:, v // 2]] = 1.0
            return L, True
        a, b = tree.left[v], tree.right[v]
        rows = slice(root_min, root_min + R) if v == tree.root else slice(0, W)
        out = None
        for c in (a, b):
            Lc, leaf = vec(c)
            M = mats[c][rows, :W]
            fac = M[:, counts[:, c // 2]].T.copy() if leaf else gemm(Lc, M)      # a leaf edge is a column gather
            out = fac if out is None else out * fac
        return out, False

    return vec(tree.root)[0]


def main():
    name = "configs[1]"
    cfg = bench_data.CONFIGS[name]
    nw = bench_data.config_tree(name)
    lam0 = bench_data.default_lambda(nw)
    counts = bench_data.config_chunk(name, 0)
    # a stratified sample: small families, the largest ones, and the most uneven ones (largest max / (min + 1) ratio)
    mx, mn = counts.max(axis=1), counts.min(axis=1)
    uneven = np.argsort(-(mx / (mn + 1.0)))[: N_FAM // 4]
    big = np.argsort(-mx)[: N_FAM // 4]
    rest = np.random.RandomState(0).choice(len(counts), N_FAM // 2, replace=False)
    idx = np.unique(np.concatenate([uneven, big, rest]))
    counts = counts[idx]
    t = oracle.parse_newick(nw)
    W, R, root_min = 251, 250, 1
    mats = oracle.node_matrices(t, [lam0] * t.n_nodes, [-1.0] * t.n_nodes, 250)
    ref = prune(t, mats, counts, W, R, root_min, lambda A, B: A @ B.T)
    ref_max = ref.max(axis=1)
    out_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r2_ozaki_error_study.jsonl")
    with open(out_path, "w") as fp:
        for S in (3, 4, 6, 8, 10, 12, 16, 20, 24):
            got = prune(t, mats, counts, W, R, root_min, lambda A, B: ozaki_gemm(A, B, S))
            gm = got.max(axis=1)
            rel = np.abs(gm - ref_max) / ref_max
            big_entries = ref > (1e-12 * ref_max)[:, None]            # root entries within 1e-12 of the row maximum
            rel_all = (np.abs(got - ref) / np.where(ref > 0, ref, 1.0))[big_entries]
            line = {"slices": S, "bits_kept": S * BITS, "int8_gemms_per_fp64_gemm": S * (S + 1) // 2, "families": int(len(counts)),
                    "max_rel_err_of_max_likelihood": float(rel.max()), "median_rel_err": float(np.median(rel)),
                    "families_worse_than_1e-11": int((rel > 1e-11).sum()), "families_worse_than_1e-6": int((rel > 1e-6).sum()),
                    "max_rel_err_of_significant_root_entries": float(rel_all.max()),
                    "fp64_equiv_tflops_at_int8_peak_4500": 4500.0 / (S * (S + 1) // 2)}
            print(json.dumps(line), flush=True)
            fp.write(json.dumps(line) + "\n")


if __name__ == "__main__":
    main()
