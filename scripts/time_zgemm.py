"""CUDA-event timing of Z = Qs^T Y (24 x 24 times 24 x (500 * 2000)) through gpcsd_dgemm."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gpcsd_b200 import _lib as L
nx, nt, N = 24, 500, 2000
Q = torch.randn(nx, nx, dtype=torch.float64, device="cuda")
Y = torch.randn(nx, nt * N, dtype=torch.float64, device="cuda")
Z = torch.empty_like(Y)
flush = torch.empty(64 << 20, dtype=torch.float64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
ts = []
for i in range(12):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    L.call("gpcsd_dgemm", 0, nx, nt * N, nx, Q.data_ptr(), nx, 0, Y.data_ptr(), nt * N, 0, Z.data_ptr(), nt * N, 0, 1, st)
    e1.record(); torch.cuda.synchronize()
    if i >= 2: ts.append(e0.elapsed_time(e1))
ms = sorted(ts)[len(ts) // 2]
print("Z gemm: %.1f us, %.2f TB/s (384 MB algorithmic), max err %.2e" % (1e3 * ms, 384e6 / (ms * 1e-3) * 1e-12, float((Z - Q @ Y).abs().max())))
