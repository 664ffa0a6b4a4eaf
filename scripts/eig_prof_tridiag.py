"""Phase timers of the tridiagonalisation kernel (build with GPCSD_NVCC_FLAGS=-DGPCSD_EIG_PROF)."""
import sys, ctypes; sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import numpy as np, torch
from gpcsd_b200 import _lib as L
lib = L.load()
st = torch.cuda.current_stream().cuda_stream
names = ["looptop", "wait", "S-sums||scalars", "coeffs+x+partials", "merged-reduce", "send", "record", "deferred-update", "keep"]
for n in (24, 125, 250):
    ld = n + (n&1); nmat = 1
    A = torch.randn(nmat,n,n,dtype=torch.float64,device="cuda"); A = A + A.transpose(1,2)
    stack = torch.zeros(nmat,n,ld,dtype=torch.float64,device="cuda"); stack[:,:,:n]=A
    d=torch.zeros(nmat,n,dtype=torch.float64,device="cuda"); e=torch.zeros_like(d); tau=torch.zeros_like(d); V=torch.zeros_like(stack)
    for _ in range(3):
        L.call("gpcsd_tridiag", n, nmat, stack.data_ptr(), ld, d.data_ptr(), e.data_ptr(), V.data_ptr(), ld, tau.data_ptr(), st)
    out = (ctypes.c_longlong*64)()
    lib.gpcsd_dbg_prof(out)
    v = np.array(list(out)[:9], dtype=float)/(n-2)
    print("n=%d cycles/column:"%n, "  ".join("%s %.0f"%(a,b) for a,b in zip(names, v)), " total %.0f"%v.sum())
