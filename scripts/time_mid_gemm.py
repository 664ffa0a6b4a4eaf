"""CUDA-event timing + check of gpcsd_dgemm on the mid-size square products of one evaluation (rotations Q C Q^T, eigenvector
products of gpcsd_eigh_dc): few 128 x 64 tiles, so the latency-mode path (32 x 32 tiles) applies."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gpcsd_b200 import _lib as L
L.load()
st = torch.cuda.current_stream().cuda_stream
for (M, N, K, batch, tb) in ((250, 250, 250, 2, 0), (250, 250, 250, 2, 1), (192, 192, 192, 2, 0), (125, 125, 125, 2, 0), (384, 384, 384, 1, 0), (250, 250, 250, 8, 0)):
    ld = lambda n: n + (n & 1)
    A = torch.randn(batch, M, ld(K), dtype=torch.float64, device="cuda")
    B = torch.randn(batch, N if tb else K, ld(K) if tb else ld(N), dtype=torch.float64, device="cuda")
    C = torch.zeros(batch, M, ld(N), dtype=torch.float64, device="cuda")
    f = lambda: L.call("gpcsd_dgemm", tb, M, N, K, A.data_ptr(), A.stride(1), A.stride(0), B.data_ptr(), B.stride(1), B.stride(0),
                       C.data_ptr(), C.stride(1), C.stride(0), batch, st)
    for _ in range(3): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(50): f()
    e1.record(); torch.cuda.synchronize()
    ref = A[:, :, :K] @ (B[:, :, :K].transpose(1, 2) if tb else B[:, :K, :N])
    err = (C[:, :, :N] - ref).abs().max().item() / ref.abs().max().item()
    print("M=%d N=%d K=%d batch=%d transB=%d: %.1f us  (max rel err %.1e)" % (M, N, K, batch, tb, 1e3 * e0.elapsed_time(e1) / 50, err))
