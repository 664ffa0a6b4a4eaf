"""Phase timers of the divide-and-conquer kernel (build with GPCSD_NVCC_FLAGS=-DGPCSD_EIG_PROF)."""
import sys, ctypes; sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import numpy as np, torch
from gpcsd_b200 import _lib as L
lib = L.load()
st = torch.cuda.current_stream().cuda_stream
names = ["init", "table+z", "sort", "deflate", "rot+sync", "secular+sync", "zhat+sync", "update+sync", "final"]
for n in (24, 125, 250):
    ld = n + (n&1); nmat = 1
    t = np.arange(n)*1.0; dd = t[:,None]-t[None,:]
    K = 0.5*np.exp(-0.5*dd**2/400.0)+0.2*np.exp(-np.abs(dd)/5.0)
    stack = torch.zeros(nmat,n,ld,dtype=torch.float64,device="cuda"); stack[:,:,:n]=torch.from_numpy(K).cuda()
    d=torch.zeros(nmat,n,dtype=torch.float64,device="cuda"); e=torch.zeros_like(d); tau=torch.zeros_like(d); V=torch.zeros_like(stack); XT=torch.zeros_like(stack); W=torch.zeros_like(d)
    nws = L.query("gpcsd_eigh_dc_ws_doubles", n, ld, nmat); ws = torch.zeros(nws,dtype=torch.float64,device="cuda")
    L.call("gpcsd_tridiag", n, nmat, stack.data_ptr(), ld, d.data_ptr(), e.data_ptr(), V.data_ptr(), ld, tau.data_ptr(), st)
    for _ in range(3):
        L.call("gpcsd_tridiag_eig", n, nmat, d.data_ptr(), e.data_ptr(), W.data_ptr(), XT.data_ptr(), ld, ws.data_ptr(), nws, 0, st)
    out = (ctypes.c_longlong*64)()
    lib.gpcsd_dbg_prof(out)
    v = np.array(list(out)[:32], dtype=float)
    print("n=%d cycles by phase:"%n, "  ".join("%s %.0f"%(a,b) for a,b in zip(names, v[:9])), " total %.0f"%v[:9].sum())
    print("     cycles by level:", " ".join("L%d %.0f"%(i, v[16+i]) for i in range(1, 9)))
    lp = (ctypes.c_longlong*256)()
    lib.gpcsd_dbg_lp(lp)
    T = np.array(list(lp), dtype=float).reshape(16, 16)
    print("     level x phase (cycles; columns: %s)" % " ".join(names[1:8]))
    for Lv in range(1, 9):
        if T[Lv, 1:8].sum() > 0:
            print("     L%d: %s" % (Lv, " ".join("%7.0f" % x for x in T[Lv, 1:8])))
