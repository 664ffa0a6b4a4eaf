"""Where does one configs[1] loglik+grad evaluation go: GPU time per ABI call vs wall clock (host/launch overhead)."""
import sys, time; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, torch
from oracle import synth
from helpers import engine_from_oracle
from gpcsd_b200 import _lib as L
x, t = synth.geometry_1d(24, 500, ms_grid=True)
rng = np.random.default_rng(2)
om = synth.model_1d(x, t, a=-200.0, b=2600.0, sig2n=1e-2 * np.exp(0.3 * rng.standard_normal(24)))
lfp = np.random.default_rng(1).standard_normal((24, 500, 2000))
eng, hp = engine_from_oracle(om, lfp)
for _ in range(5): eng.loglik_grad(hp)
torch.cuda.synchronize()
N = 20
t0 = time.perf_counter()
for _ in range(N): eng.loglik_grad(hp)
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / N
names = list(L.SIGNATURES.keys())
eng.timers = {n: [] for n in names}
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
