import torch, sys
m, n, k = 4096, 32768, 4096
a = torch.randn(m, k, dtype=torch.float64, device="cuda")
b = torch.randn(n, k, dtype=torch.float64, device="cuda")
for _ in range(2):
    c = a @ b.t()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    c = a @ b.t()
e1.record(); torch.cuda.synchronize()
print("cuBLAS NT %dx%dx%d: %.2f TFLOP/s" % (m, n, k, 2 * m * n * k * 5 / (e0.elapsed_time(e1) * 1e-3) / 1e12))
