"""Measures the FP64 GEMM rate of the box (cuBLAS DGEMM through torch.matmul, 8192^3), the denominator of the
FP64 roofline that MEASURED_PEAKS.json does not carry. Writes gpurun_out/fp64_peak.json."""
import json
import os
import time

import torch

n = 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda")
b = torch.randn(n, n, dtype=torch.float64, device="cuda")
for _ in range(2):
    (a @ b)
torch.cuda.synchronize()
best = 0.0
times = []
for _ in range(6):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    c = a @ b
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e)
    times.append(ms)
    best = max(best, 2 * n ** 3 / (ms * 1e-3) / 1e12)
# sustained: back to back for ~3 s
t0 = time.time()
cnt = 0
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
while time.time() - t0 < 3.0:
    c = a @ b
    cnt += 1
e.record()
torch.cuda.synchronize()
sus = cnt * 2 * n ** 3 / (s.elapsed_time(e) * 1e-3) / 1e12
out = {"fp64_tflops": best, "fp64_tflops_sustained": sus, "how": "torch.matmul float64 8192^3 (cuBLAS DGEMM), best of 6 / 3 s loop",
       "gpu": torch.cuda.get_device_name(0), "times_ms": times}
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/fp64_peak.json", "w"), indent=1)
print(json.dumps(out))
