"""CUDA-event timing of gpcsd_eigh_dc on small orders (Jacobi path) for 1 and 192 stacked GP matrices."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gpcsd_b200 import _lib as L
for n, nmat in ((24, 1), (24, 64), (25, 128), (32, 192)):
    x = np.linspace(0, 2300, n); dd = x[:, None] - x[None, :]
    K = np.exp(-0.5 * dd ** 2 / 200.0 ** 2) + 1e-8 * np.eye(n)
    ld = (n + 1) // 2 * 2
    st = torch.zeros((nmat, n, ld), dtype=torch.float64, device="cuda"); st[:, :, :n] = torch.from_numpy(K).cuda()
    QT = torch.zeros_like(st); W = torch.zeros((nmat, n), dtype=torch.float64, device="cuda")
    info = torch.zeros(nmat, dtype=torch.int32, device="cuda")
    ws = torch.zeros(max(L.query("gpcsd_eigh_dc_ws_doubles", n, ld, nmat), 1), dtype=torch.float64, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    ts = []
    for i in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.call("gpcsd_eigh_dc", n, nmat, st.data_ptr(), ld, QT.data_ptr(), ld, W.data_ptr(), ws.data_ptr(), ws.numel(), info.data_ptr(), s)
        e1.record(); torch.cuda.synchronize()
        if i >= 2: ts.append(e0.elapsed_time(e1))
    lam = np.linalg.eigvalsh(K)
    Wh = W[0].cpu().numpy(); Q = QT[0, :, :n].cpu().numpy().T
    print("n=%d nmat=%d: %.1f us; max rel eig err %.1e, orth %.1e, resid %.1e" % (n, nmat, 1e3 * np.median(ts),
          np.max(np.abs(Wh - lam) / np.abs(lam)), np.max(np.abs(Q.T @ Q - np.eye(n))), np.max(np.abs(K @ Q - Q * Wh))))
