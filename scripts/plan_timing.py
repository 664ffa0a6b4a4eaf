"""Time one model's loglik+grad through the native plan (graph on / off) and through the call-by-call Python path."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gpcsd_b200.engine import HyperParams, KronEngine
from gpcsd_b200.covariances import GPCSD1DSpatialCovSE, GPCSD2DSpatialCovSE

def timeit(fn, n):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return 1e3 * (time.perf_counter() - t0) / n

def case(name):
    rng = np.random.default_rng(0)
    if name == "cfg1":
        x = np.linspace(0, 2300, 24)[:, None]; t = np.arange(500.0)[:, None]
        sc = GPCSD1DSpatialCovSE(x, a=-200.0, b=2600.0, ngl=100)
        eng = KronEngine(1, x, t, dict(gl_x=sc.gl_x, gl_w=sc.gl_w))
        eng.set_lfp(torch.randn(24, 500, 2000, dtype=torch.float64, device="cuda"))
        hp = lambda: HyperParams(R=100.0 * np.exp(0.05 * rng.standard_normal()), ells=(200.0,), temporal=[(0, 20.0, 0.5 / 300), (1, 5.0, 0.7 / 300)], sig2n=1e-2 * np.exp(0.3 * rng.standard_normal(24)))
    elif name == "cfg0":
        x = np.linspace(0, 2300, 24)[:, None]; t = np.linspace(0, 50.0, 50)[:, None]
        sc = GPCSD1DSpatialCovSE(x, a=0.0, b=2300.0, ngl=100)
        eng = KronEngine(1, x, t, dict(gl_x=sc.gl_x, gl_w=sc.gl_w))
        eng.set_lfp(torch.randn(24, 50, 50, dtype=torch.float64, device="cuda"))
        hp = lambda: HyperParams(R=100.0 * np.exp(0.05 * rng.standard_normal()), ells=(200.0,), temporal=[(0, 20.0, 0.5), (1, 5.0, 0.7)], sig2n=1e-2)
    else:
        ch = np.arange(384); X = np.stack([np.array([16.0, 48.0, 0.0, 32.0])[ch % 4], 20.0 * np.floor(ch / 2)], axis=1)
        t = (0.4 * np.arange(250.0))[:, None]
        sc = GPCSD2DSpatialCovSE(X, a1=-16.0, b1=64.0, a2=-100.0, b2=float(X[:, 1].max()) + 100.0, ngl1=30, ngl2=120)
        eng = KronEngine(2, X, t, dict(gl_x1=sc.gl_x1, gl_w1=sc.gl_w1, gl_x2=sc.gl_x2, gl_w2=sc.gl_w2))
        eng.set_lfp(torch.randn(384, 250, 500, dtype=torch.float64, device="cuda"))
        hp = lambda: HyperParams(R=100.0 * np.exp(0.05 * rng.standard_normal()), ells=(40.0, 200.0), temporal=[(0, 5.0, 0.5 / 300), (1, 1.0, 0.7 / 300)], sig2n=0.5, eps=1.0)
    return eng, hp

for name in sys.argv[1:] or ["cfg0", "cfg1", "cfg2"]:
    eng, hp = case(name)
    n = 200 if name == "cfg0" else 20
    eng.use_plan = False
    t_step = timeit(lambda: eng.loglik_grad(hp()), n)
    eng.use_plan = True
    t_plan = timeit(lambda: eng.loglik_grad(hp()), n)
    pl = list(eng._plans.values())[0]
    pl.set_graph(False)
    t_eager = timeit(lambda: eng.loglik_grad(hp()), n)
    pl.set_graph(True)
    t_plan2 = timeit(lambda: eng.loglik_grad(hp()), n)
    print("%s: stepwise %.3f ms | plan graph %.3f ms | plan eager %.3f ms | plan graph again %.3f ms | launches %d" % (name, t_step, t_plan, t_eager, t_plan2, pl.last_launches()), flush=True)
