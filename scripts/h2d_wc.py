"""Pinned host->device bandwidth: default pinned allocation vs write-combined (cudaHostAllocWriteCombined)."""
import ctypes, torch
rt = ctypes.CDLL("libcudart.so.12")
n = 192_000_000
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
def alloc(flags):
    p = ctypes.c_void_p()
    assert rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(n), ctypes.c_uint(flags)) == 0
    ctypes.memset(p, 1, n)
    return p
for name, flags in (("default pinned", 0), ("write-combined", 4), ("portable+mapped", 3)):
    p = alloc(flags)
    st = torch.cuda.current_stream().cuda_stream
    best = 0.0
    for rep in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4):
            assert rt.cudaMemcpyAsync(ctypes.c_void_p(dev.data_ptr()), p, ctypes.c_size_t(n), ctypes.c_int(1), ctypes.c_void_p(st)) == 0
        e1.record(); torch.cuda.synchronize()
        best = max(best, 4 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    print("%s: %.1f GB/s" % (name, best))
    rt.cudaFreeHost(p)
