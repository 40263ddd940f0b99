"""Measure the FP64 roofline denominators on this box: cuBLAS DGEMM (torch.matmul f64)
burst and sustained.  Writes profiles/fp64_peak.json.  Library GEMM = reference point only."""
import json, os, sys, time
import torch
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda")
b = torch.randn(n, n, dtype=torch.float64, device="cuda")
for _ in range(3):
    c = a @ b
torch.cuda.synchronize()
best = 0.0
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
    best = max(best, 2 * n**3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); k = 0; t0 = time.time()
while time.time() - t0 < 3.0:
    c = a @ b; k += 1
    if k % 8 == 0: torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
sus = 2 * n**3 * k / (e0.elapsed_time(e1) * 1e-3) / 1e12
out = {"fp64_dgemm_tflops_burst": best, "fp64_dgemm_tflops_sustained": sus, "n": n,
       "how": "torch.matmul float64 8192^3 (cuBLAS DGEMM), CUDA events, best of 10 / 3 s loop",
       "gpu": torch.cuda.get_device_name(0)}
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/fp64_peak.json", "w"), indent=1)
