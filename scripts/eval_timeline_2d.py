"""configs[2] (Neuropixels 384 x 250 x 500): GPU time per ABI call vs wall clock."""
import sys, time; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, torch
from oracle import synth
from helpers import engine_from_oracle
from gpcsd_b200 import _lib as L
X, t = synth.geometry_neuropixels(384, 250, 0.4)
om = synth.model_2d(X, t, ngl1=30, ngl2=120, a1=-16.0, b1=64.0, a2=-100.0, b2=float(X[:, 1].max()) + 100.0, eps=1.0, sig2n=0.5)
lfp = np.random.default_rng(1).standard_normal((384, 250, 500))
eng, hp = engine_from_oracle(om, lfp)
for _ in range(5): eng.loglik_grad(hp)
torch.cuda.synchronize()
N = 10
t0 = time.perf_counter()
for _ in range(N): eng.loglik_grad(hp)
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / N
eng.timers = {n: [] for n in L.SIGNATURES.keys()}
for _ in range(N): eng.loglik_grad(hp)
torch.cuda.synchronize()
tot = 0.0; rows = []
for n, lst in eng.timers.items():
    if lst:
        ms = sum(a.elapsed_time(b) for a, b in lst) / N
        rows.append((ms, n, len(lst) / N)); tot += ms
eng.timers = None
for ms, n, c in sorted(rows, reverse=True): print("  %-28s %6.3f ms  (%.0f calls/eval)" % (n, ms, c))
print("wall %.3f ms/eval; sum of per-call GPU spans %.3f ms; calls/eval %d" % (wall * 1e3, tot, sum(c for _, _, c in rows)))
