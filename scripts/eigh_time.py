"""Developer measurement script (run from the repo root on a B200); numbers quoted in profiles/r01d_eigensolver.md / DESIGN.md."""
import sys; sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import numpy as np, torch
from gpcsd_b200 import _lib as L
L.load()
st = torch.cuda.current_stream().cuda_stream
def run(f, reps=20):
    for _ in range(3): f()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True); e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/reps
for n, nmat in ((250,2),(250,1),(192,2),(125,2),(50,2),(24,2),(24,64)):
    ld = n + (n&1)
    t = np.arange(n)*1.0; dd = t[:,None]-t[None,:]
    K = 0.5*np.exp(-0.5*dd**2/400.0)+0.2*np.exp(-np.abs(dd)/5.0)
    stack = torch.zeros(nmat,n,ld,dtype=torch.float64,device="cuda"); stack[:,:,:n]=torch.from_numpy(K).cuda()
    d=torch.zeros(nmat,n,dtype=torch.float64,device="cuda"); e=torch.zeros_like(d); tau=torch.zeros_like(d); V=torch.zeros_like(stack); XT=torch.zeros_like(stack); W=torch.zeros_like(d)
    nws = L.query("gpcsd_eigh_dc_ws_doubles", n, ld, nmat); ws = torch.zeros(nws,dtype=torch.float64,device="cuda")
    t1 = run(lambda: L.call("gpcsd_tridiag", n, nmat, stack.data_ptr(), ld, d.data_ptr(), e.data_ptr(), V.data_ptr(), ld, tau.data_ptr(), st))
    t2 = run(lambda: L.call("gpcsd_tridiag_eig", n, nmat, d.data_ptr(), e.data_ptr(), W.data_ptr(), XT.data_ptr(), ld, ws.data_ptr(), nws, 0, st))
    t3 = run(lambda: L.call("gpcsd_backtransform", n, nmat, V.data_ptr(), ld, tau.data_ptr(), XT.data_ptr(), ld, st))
    t4 = run(lambda: L.call("gpcsd_eigh_dc", n, nmat, stack.data_ptr(), ld, XT.data_ptr(), ld, W.data_ptr(), ws.data_ptr(), nws, 0, st))
    lam = np.linalg.eigvalsh(K); err = np.max(np.abs(W[0].cpu().numpy()-lam))/lam.max()
    print("n=%d nmat=%d: tridiag %.3f  dc %.3f  backtransform %.3f  | eigh_dc %.3f ms  (eig err %.1e)"%(n,nmat,t1,t2,t3,t4,err), flush=True)
