"""cuBLAS DGEMM peak (fp64 roofline denominator) — burst and sustained. Prints one JSON line."""
import json, time, torch
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda")
b = torch.randn(n, n, dtype=torch.float64, device="cuda")
c = torch.empty_like(a)
for _ in range(3):
    torch.matmul(a, b, out=c)
torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); torch.matmul(a, b, out=c); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
burst = 2 * n**3 / best * 1e-9
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
reps = 40
e0.record()
for _ in range(reps):
    torch.matmul(a, b, out=c)
e1.record(); torch.cuda.synchronize()
sus = 2 * n**3 * reps / e0.elapsed_time(e1) * 1e-9
print(json.dumps({"dgemm_n": n, "fp64_tflops_burst": burst, "fp64_tflops_sustained": sus, "burst_ms": best}))
