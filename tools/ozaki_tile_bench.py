"""Drive tools/ozaki_tile_bench.cu with REAL operands of BASELINE configs[1]: A = the leaf-pair vectors of the first cherry of
the bench tree for all 50 k families (products of two gathered matrix columns, cafe_tree.c:204-210 / :266-270; they span hundreds
of decades), B = the transition matrix of the cherry's own branch (K1's output).  Reports, per slice count S, the time of the
slicing pass and of the tcgen05 kernel, fp64-equivalent TFLOP/s, and the error of Out = A B^T against an 80-bit reference on a
row sample — next to plain fp64 (numpy) and to cuBLAS DGEMM on the same shape.  Prints one JSON line.

    python tools/ozaki_tile_bench.py            # needs a B200; builds tools/libozaki_tile.so if missing
"""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cafe_b200 import gpu as cgpu, host as chost, synth  # noqa: E402

LIB = os.path.join(ROOT, "tools", "libozaki_tile.so")
if not os.path.exists(LIB):
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
                    "-o", LIB, os.path.join(ROOT, "tools", "ozaki_tile_bench.cu")], check=True)
L = C.CDLL(LIB)
dp = C.POINTER(C.c_double)

F, T, MS = 50000, 20, 200
newick = synth.random_tree(T, 1)
counts, lam0 = synth.simulate_table(newick, F, MS, seed=10)
tree = chost.parse_tree(newick)
rg = chost.init_family_size(MS)
ranges = (rg["min"], rg["max"], rg["root_min"], rg["root_max"])
n = tree.n_nodes
g = cgpu.CafeGpu(0)
g.set_tree(tree.left, tree.right, tree.branchlength); g.set_ranges(*ranges)
g.set_lnc_table(chost.lnc_table(max(ranges[1], ranges[3])))
g.set_rates(np.full(n, lam0), np.full(n, -1.0)); g.build_matrices()
W = ranges[1] + 1
left, right = tree.left, tree.right
cherry = next(v for v in range(n) if left[v] >= 0 and left[left[v]] < 0 and left[right[v]] < 0)
Ma, Mb, Mv = g.get_matrix(int(left[cherry])), g.get_matrix(int(right[cherry])), g.get_matrix(cherry)
g.close()
ca, cb = counts[:, left[cherry] // 2], counts[:, right[cherry] // 2]
K = 256
A = np.zeros((F, K)); B = np.zeros((K, K))
A[:, :W] = Ma[:W, ca].T * Mb[:W, cb].T          # L[f][j] = M_a[j][c_a] * M_b[j][c_b]
B[:W, :W] = Mv[:W, :W]                           # Out[f][i] = sum_j M_v[i][j] L[f][j]
Fp = (F + 127) // 128 * 128
A = np.ascontiguousarray(np.vstack([A, np.zeros((Fp - F, K))]))
pos = A[A > 0]
info = {"F": int(Fp), "N": K, "K": K, "W": int(W), "decades_A": float(np.log10(pos.max()) - np.log10(pos.min())),
        "decades_per_row_median": float(np.median([np.log10(r[r > 0].max()) - np.log10(r[r > 0].min()) for r in A[:2000] if (r > 0).any()]))}

rs = np.random.RandomState(3)
sample = np.sort(rs.choice(F, 256, replace=False))
ref = (A[sample].astype(np.longdouble) @ B.T.astype(np.longdouble))          # 64-bit mantissa reference
ref64 = ref.astype(np.float64)


def errors(out):
    o = out[sample].astype(np.longdouble)
    rowmax = ref.max(axis=1, keepdims=True)
    with np.errstate(divide="ignore", invalid="ignore"):
        rel = np.abs(o - ref) / ref
    rel = np.where(ref > 0, rel, 0.0)
    am = ref.argmax(axis=1)
    e_max = rel[np.arange(len(sample)), am]
    near = ref >= rowmax * 1e-12
    return {"max_rel_err_of_row_max": float(e_max.max()), "max_rel_err_entries_within_1e-12_of_row_max": float(rel[near].max()),
            "max_rel_err_all_positive_entries": float(rel.max())}


flops = 2.0 * Fp * K * K
res = {"operands": info, "flops_per_gemm": flops, "fp64_numpy": errors(A @ B.T)}
out = np.zeros((Fp, K)); ms = (C.c_float * 2)()
for S in (4, 6, 7, 8):
    per = {}
    for mode, name in ((0, "full"), (1, "mma_and_feed_only"), (2, "feed_and_epilogue_only")):
        rc = L.ozaki_tile_gemm(A.ctypes.data_as(dp), B.ctypes.data_as(dp), Fp, K, S, out.ctypes.data_as(dp), ms, 10, mode)
        if rc:
            per[name] = {"error": rc}; continue
        per[name] = {"slice_ms": ms[0], "kernel_ms": ms[1]}
        if mode == 0:
            per["errors"] = errors(out)
            per["int8_mmas_per_fp64_gemm"] = S * (S + 1) // 2
            per["fp64_equiv_tflops_kernel"] = flops / (ms[1] * 1e-3) * 1e-12
            per["fp64_equiv_tflops_with_slicing"] = flops / ((ms[0] + ms[1]) * 1e-3) * 1e-12
            per["int8_tops_kernel"] = flops * (S * (S + 1) // 2) / (ms[1] * 1e-3) * 1e-12
    res[f"S={S}"] = per

try:
    import torch
    a = torch.from_numpy(A).cuda(); b = torch.from_numpy(B).cuda(); c = torch.empty((Fp, K), dtype=torch.float64, device="cuda")
    for _ in range(3):
        torch.matmul(a, b.T, out=c)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        torch.matmul(a, b.T, out=c)
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 10
    res["cublas_dgemm_same_shape"] = {"ms": t, "tflops": flops / (t * 1e-3) * 1e-12, "errors": errors(c.cpu().numpy())}
except Exception as e:  # noqa: BLE001
    res["cublas_dgemm_same_shape"] = {"error": repr(e)}
res["k2_dmma_kernel_tflops_configs1"] = 31.7
print(json.dumps(res))
