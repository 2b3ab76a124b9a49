"""Measures the sustained cuBLAS DGEMM rate (the FP64 roofline denominator) and, for context,
torch/cuSOLVER Cholesky at the cfg-4 size.  Not part of the product path."""
import json, sys, time
import torch

def dgemm_peak(n=8192, burst_reps=5, sustain_s=3.0):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    c = torch.empty_like(a)
    for _ in range(2):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(burst_reps):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); torch.matmul(a, b, out=c); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    burst = 2 * n ** 3 / best * 1e-9
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    reps = max(3, int(sustain_s * 1e3 / best))
    e0.record()
    for _ in range(reps):
        torch.matmul(a, b, out=c)
    e1.record(); e1.synchronize()
    sustained = 2 * n ** 3 * reps / e0.elapsed_time(e1) * 1e-9
    return burst, sustained

if __name__ == "__main__":
    burst, sus = dgemm_peak()
    out = {"dgemm_tflops_burst": burst, "dgemm_tflops_sustained": sus}
    n = 16384
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    spd = a @ a.T / n + torch.eye(n, dtype=torch.float64, device="cuda") * 4
    del a
    torch.linalg.cholesky(spd); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); torch.linalg.cholesky(spd); e1.record(); e1.synchronize()
    ms = e0.elapsed_time(e1)
    out["cusolver_potrf_16384_ms"] = ms
    out["cusolver_potrf_16384_tflops"] = n ** 3 / 3 / ms * 1e-9
    print(json.dumps(out))
