"""Feasibility probe: a green context holding all SMs but 16 for the GEMM-phase stream.  Does a persistent GEMM launched
there leave GPC-aligned SMs for the 8-CTA eigensolver clusters of another stream?  (eigh latency alone / under a GEMM on a
normal stream / under a GEMM on the green-context stream; stream capture on the green-context stream.)"""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cuda.bindings import driver as cu
from gpcsd_b200 import _lib as L
L.load()
torch.cuda.init(); torch.zeros(1, device="cuda")

def ck(r):
    if isinstance(r, tuple):
        err, rest = r[0], r[1:]
    else:
        err, rest = r, ()
    assert err == cu.CUresult.CUDA_SUCCESS, err
    return rest[0] if len(rest) == 1 else rest

dev = ck(cu.cuDeviceGet(0))
res = ck(cu.cuDeviceGetDevResource(dev, cu.CUdevResourceType.CU_DEV_RESOURCE_TYPE_SM))
print("device SMs:", res.sm.smCount)
result, nb, remaining = ck(cu.cuDevSmResourceSplitByCount(1, res, 0, 16))
print("split: group", result[0].sm.smCount, "remaining", remaining.sm.smCount, "groups", nb)
desc = ck(cu.cuDevResourceGenerateDesc([remaining], 1))
gctx = ck(cu.cuGreenCtxCreate(desc, dev, cu.CUgreenCtxCreate_flags.CU_GREEN_CTX_DEFAULT_STREAM))
gstream = ck(cu.cuGreenCtxStreamCreate(gctx, cu.CUstream_flags.CU_STREAM_NON_BLOCKING, 0))
gs = int(gstream)
print("green stream handle", hex(gs))

# a long persistent GEMM (projection-like): M=250, N=2000, K=250, batch 96 -> ~1 ms
M, N, K, B = 250, 2000, 250, 96
A = torch.randn(1, M, K, dtype=torch.float64, device="cuda")
Bm = torch.randn(B, K, N, dtype=torch.float64, device="cuda")
C = torch.zeros(B, M, N, dtype=torch.float64, device="cuda")
def gemm(stream):
    L.call("gpcsd_dgemm", 0, M, N, K, A.data_ptr(), K, 0, Bm.data_ptr(), N, K * N, C.data_ptr(), N, M * N, B, stream)
# eigensolver problem: two 250-order matrices
n = 250; ld = n
t = np.arange(n) * 1.0; dd = t[:, None] - t[None, :]
Kt = 0.5 * np.exp(-0.5 * dd ** 2 / 400.0) + 0.2 * np.exp(-np.abs(dd) / 5.0)
st_ = torch.zeros(2, n, ld, dtype=torch.float64, device="cuda"); st_[:, :, :n] = torch.from_numpy(Kt).cuda()
QT = torch.zeros_like(st_); W = torch.zeros(2, n, dtype=torch.float64, device="cuda")
ws = torch.zeros(L.query("gpcsd_eigh_dc_ws_doubles", n, ld, 2), dtype=torch.float64, device="cuda")
es = torch.cuda.Stream()
def eigh():
    L.call("gpcsd_eigh_dc", n, 2, st_.data_ptr(), ld, QT.data_ptr(), ld, W.data_ptr(), ws.data_ptr(), ws.numel(), 0, es.cuda_stream)

def timed(fn, stream_obj, reps=5):
    out = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream_obj); fn(); e1.record(stream_obj); torch.cuda.synchronize(); out.append(e0.elapsed_time(e1))
    return float(np.median(out))

ns = torch.cuda.Stream()
gemm(ns.cuda_stream); gemm(gs); eigh(); torch.cuda.synchronize()
ref = (A[0] @ Bm[3]).cpu(); print("gemm on green stream correct:", float((C[3].cpu() - ref).abs().max()) < 1e-9)
print("eigh alone: %.3f ms" % timed(eigh, es))
print("gemm alone on normal stream: %.3f ms" % timed(lambda: gemm(ns.cuda_stream), ns))
gev0, gev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def gemm_green_timed():
    # events cannot be recorded by torch on a foreign stream object: bracket with a normal stream waiting on nothing; use wall sync
    import time
    torch.cuda.synchronize(); t0 = time.perf_counter(); gemm(gs); torch.cuda.synchronize(); return 1e3 * (time.perf_counter() - t0)
print("gemm alone on green stream (wall): %.3f ms" % np.median([gemm_green_timed() for _ in range(5)]))
def under(stream_handle):
    out = []
    for _ in range(5):
        torch.cuda.synchronize()
        for _ in range(3): gemm(stream_handle)          # ~3 ms of persistent GEMMs in flight
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(es); eigh(); e1.record(es); torch.cuda.synchronize(); out.append(e0.elapsed_time(e1))
    return float(np.median(out))
print("eigh under GEMMs on a normal stream: %.3f ms" % under(ns.cuda_stream))
print("eigh under GEMMs on the green stream: %.3f ms" % under(gs))
# stream capture on the green stream
rt = ctypes.CDLL("libcudart.so.12")
g = ctypes.c_void_p()
r1 = rt.cudaStreamBeginCapture(ctypes.c_void_p(gs), 2)
gemm(gs)
r2 = rt.cudaStreamEndCapture(ctypes.c_void_p(gs), ctypes.byref(g))
print("capture on green stream: begin rc %d end rc %d graph %s" % (r1, r2, bool(g.value)))
if r2 == 0 and g.value:
    ge = ctypes.c_void_p()
    r3 = rt.cudaGraphInstantiate(ctypes.byref(ge), g, 0)
    r4 = rt.cudaGraphLaunch(ge, ctypes.c_void_p(gs)) if r3 == 0 else -1
    torch.cuda.synchronize()
    print("instantiate rc %d launch rc %d; eigh under graph-launched GEMMs:" % (r3, r4))
    out = []
    for _ in range(5):
        torch.cuda.synchronize()
        for _ in range(3): rt.cudaGraphLaunch(ge, ctypes.c_void_p(gs))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(es); eigh(); e1.record(es); torch.cuda.synchronize(); out.append(e0.elapsed_time(e1))
    print("  %.3f ms" % float(np.median(out)))
