"""FP64 peak of this pool's B200 by the protocol of MEASURED_PEAKS.json / BASELINE.md section 3:
cuBLAS DGEMM 8192^3 through torch.matmul (float64): best of 10 (burst) and back to back for 4 s (sustained),
CUDA events; plus an FP64 STREAM-style copy for reference.  Prints one JSON line.
usage: python profiles/fp64_peak.py"""
import json
import time

import torch

N = 8192
dev = torch.device("cuda:0")
a = torch.randn(N, N, dtype=torch.float64, device=dev)
b = torch.randn(N, N, dtype=torch.float64, device=dev)
c = torch.empty(N, N, dtype=torch.float64, device=dev)
flop = 2.0 * N ** 3
for _ in range(3):
    torch.matmul(a, b, out=c)
torch.cuda.synchronize()
burst = []
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    torch.matmul(a, b, out=c)
    e1.record()
    torch.cuda.synchronize()
    burst.append(e0.elapsed_time(e1) * 1e-3)
# sustained: back to back for 4 s
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
n = 0
e0.record()
while time.perf_counter() - t0 < 4.0:
    for _ in range(4):
        torch.matmul(a, b, out=c)
    n += 4
    torch.cuda.synchronize()
e1.record()
torch.cuda.synchronize()
sustained = flop * n / (e0.elapsed_time(e1) * 1e-3)
print(json.dumps({
    "gpu_name": torch.cuda.get_device_name(0),
    "fp64_dgemm_tflops": round(flop / min(burst) / 1e12, 2),
    "fp64_dgemm_tflops_sustained": round(sustained / 1e12, 2),
    "burst_ms": [round(t * 1e3, 2) for t in burst],
    "sustained_gemms": n,
    "how": "torch.matmul float64 8192^3 (cuBLAS DGEMM, 2*N^3 flop): best of 10 (burst) and back to back for 4 s (sustained), CUDA events",
}))
