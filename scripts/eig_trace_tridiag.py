import sys, ctypes; sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import numpy as np, torch
from gpcsd_b200 import _lib as L
lib = L.load()
st = torch.cuda.current_stream().cuda_stream
names = ["top", "wait", "pvsum", "x+part", "reduce", "scalars", "send", "update", "reflect"]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
ld = n + (n&1); nmat = 1
A = torch.randn(nmat,n,n,dtype=torch.float64,device="cuda"); A = A + A.transpose(1,2)
stack = torch.zeros(nmat,n,ld,dtype=torch.float64,device="cuda"); stack[:,:,:n]=A
d=torch.zeros(nmat,n,dtype=torch.float64,device="cuda"); e=torch.zeros_like(d); tau=torch.zeros_like(d); V=torch.zeros_like(stack)
for _ in range(3):
    L.call("gpcsd_tridiag", n, nmat, stack.data_ptr(), ld, d.data_ptr(), e.data_ptr(), V.data_ptr(), ld, tau.data_ptr(), st)
out = (ctypes.c_longlong*(8*16*12))()
lib.gpcsd_dbg_trace(out)
T = np.array(list(out), dtype=np.int64).reshape(8,16,12)
t0 = T[0,:3,0].min()
nw = min(8, (n+7)//8)
for kk in range(8):
    k = kk+4
    for w in range(nw):
        rows = [w*8, (w+8)*8, (w+16)*8, (w+24)*8]
        print("k=%d warp %d (rows %s)%s: "%(k, w, rows, " OWNER(k+1)" if (k+1) in rows else ""), " ".join("%s@%d"%(nm, T[kk,w,i]-t0) for i,nm in enumerate(names)))
