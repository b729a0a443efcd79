"""Measure cuBLAS DGEMM throughput (torch.matmul, float64) on the current GPU: the FP64 roofline
denominator that MEASURED_PEAKS.json does not carry.  Prints one JSON line."""
import json, sys, torch

def bench(m, n, k, reps=5):
    a = torch.randn(m, k, dtype=torch.float64, device="cuda")
    b = torch.randn(k, n, dtype=torch.float64, device="cuda")
    for _ in range(2):
        (a @ b)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2.0 * m * n * k / (best * 1e-3) / 1e12, best

if __name__ == "__main__":
    out = {"gpu": torch.cuda.get_device_name(0)}
    for (m, n, k) in [(4096, 4096, 4096), (8192, 8192, 8192), (10000, 10000, 100), (10000, 100, 10000), (5050, 5050, 100), (5050, 100, 5050), (23104, 23104, 152)]:
        tf, ms = bench(m, n, k)
        out[f"dgemm_{m}x{n}x{k}"] = {"tflops": round(tf, 2), "ms": round(ms, 4)}
    # sustained: back to back for ~2 s
    a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda"); b = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    nrep = 60
    for _ in range(nrep): a @ b
    e1.record(); torch.cuda.synchronize()
    out["dgemm_8192_sustained_tflops"] = round(2.0 * 8192**3 * nrep / (e0.elapsed_time(e1) * 1e-3) / 1e12, 2)
    print(json.dumps(out))
