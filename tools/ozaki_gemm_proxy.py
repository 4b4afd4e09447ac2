"""Throughput proxy for the split-precision (Ozaki, int8 tensor core) route on ONE edge GEMM of BASELINE configs[1]
(50 k families x 256 sizes times a 256 x 256 transition matrix), run on the GPU with LIBRARY kernels only:
torch._int_mm (cuBLASLt int8 -> int32, the tcgen05 `kind::i8` path on sm_100) for the slice products, torch element-wise
kernels for slicing and recombination.  It answers one question before anyone hand-writes the tcgen05 kernel: with the
S = 8 slices the error study asks for (tools/ozaki_error_study.py: 36 int8 GEMMs per fp64 GEMM for <= 1e-11), does the
tensor-core time alone leave room to beat the DMMA kernel (30.5 TFLOP/s achieved, 35.9 TFLOP/s cuBLAS DGEMM)?

Prints one JSON line (profiles/r2_ozaki_gemm_proxy.json).  Not part of the product path."""
import json
import sys

import torch

M, N, K, S, BITS = 50048, 256, 256, 8, 7
dev = torch.device("cuda")
torch.manual_seed(0)


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return out, best


def slice_rows(X):
    """[rows][K] fp64 >= 0 -> (S int8 planes, row scale): X ~= sum_s plane_s * scale * 2^(-BITS (s+1))."""
    mx = X.max(dim=1).values
    e = torch.where(mx > 0, torch.ceil(torch.log2(torch.where(mx > 0, mx, torch.ones_like(mx)))), torch.zeros_like(mx))
    scale = torch.ldexp(torch.ones_like(mx), e.to(torch.int32))
    rem = X / scale[:, None]
    planes = []
    for _ in range(S):
        rem = rem * (1 << BITS)
        q = torch.floor(rem)
        planes.append(q.to(torch.int8))
        rem = rem - q
    return planes, scale


# operands with the dynamic range of real node vectors: every row spans ~60 decades around a moving peak
j = torch.arange(K, device=dev, dtype=torch.float64)
peak = torch.randint(0, K, (M,), device=dev).double()
A = torch.exp(-0.5 * ((j[None, :] - peak[:, None]) / 3.0) ** 2) * torch.rand(M, K, device=dev, dtype=torch.float64)
i = torch.arange(N, device=dev, dtype=torch.float64)
B = torch.exp(-0.5 * ((j[None, :] - i[:, None]) / 4.0) ** 2) * (0.5 + torch.rand(N, K, device=dev, dtype=torch.float64))

(Ap, sa), t_slice_a = timed(lambda: slice_rows(A))
(Bp, sb), t_slice_b = timed(lambda: slice_rows(B))
BpT = [b.t().contiguous() for b in Bp]


def gemms():
    acc = [None] * S                                   # one int32 accumulator per weight p + q
    for p in range(S):
        for q in range(S - p):
            r = torch._int_mm(Ap[p], BpT[q])
            acc[p + q] = r if acc[p + q] is None else acc[p + q] + r
    return acc


acc, t_gemm = timed(gemms)
_, t_gemm_only = timed(lambda: [torch._int_mm(Ap[p], BpT[q]) for p in range(S) for q in range(S - p)])


def recombine():
    C = torch.zeros(M, N, device=dev, dtype=torch.float64)
    for w in range(S):
        C += acc[w].double() * 2.0 ** (-BITS * (w + 2))
    return C * sa[:, None] * sb[None, :]


C, t_rec = timed(recombine)
ref, t_dgemm = timed(lambda: A @ B.t())
sig = ref > 1e-9 * ref.max(dim=1, keepdim=True).values
rel = ((C - ref).abs() / ref)[sig].max().item()
flops = 2.0 * M * N * K
n_gemms = S * (S + 1) // 2
line = {"shape": [M, N, K], "slices": S, "int8_gemms": n_gemms,
        "ms": {"slice_A": t_slice_a, "slice_B": t_slice_b, "int8_gemms_only": t_gemm_only, "int8_gemms_plus_int32_adds": t_gemm,
               "recombine": t_rec, "cublas_dgemm_same_shape": t_dgemm},
        "int8_tops_achieved": n_gemms * flops / (t_gemm_only * 1e-3) * 1e-12,
        "fp64_equiv_tflops_tensor_time_only": flops / (t_gemm_only * 1e-3) * 1e-12,
        "fp64_equiv_tflops_with_library_slicing_and_recombination": flops / ((t_slice_a + t_gemm + t_rec) * 1e-3) * 1e-12,
        "dgemm_tflops_same_shape": flops / (t_dgemm * 1e-3) * 1e-12,
        "max_rel_err_entries_within_1e-9_of_row_max": rel,
        "note": "library kernels only (cuBLASLt int8, torch element-wise): an upper bound on what slicing/recombination cost when NOT fused, "
                "and a measured figure for the int8 tensor-core time of the 36 slice products"}
print(json.dumps(line))
