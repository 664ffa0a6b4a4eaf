"""Lock-step batched projected L-BFGS for many independent small problems whose objective is evaluated in ONE batched call.

Two callers:
  * fit(): the multi-start restarts of gpcsd1d.py:193-220 / gpcsd2d.py:223-260 -- the reference runs scipy's L-BFGS-B once per
    restart, sequentially; here all restarts advance together so that every step is one restart-batched loglik+grad launch
    sequence (engine / gpcsd_plan) instead of n_restarts separate ones;
  * the per-trial evoked-shift fits of auditory_lfp/fit_mean_function.py:323-328 (one scipy L-BFGS-B per trial on joblib
    workers in the reference): all trials advance together, one batched device evaluation per step.

Algorithm (per problem, vectorised over the batch): limited-memory BFGS two-loop recursion on the free variables (variables
sitting on a bound with the gradient pointing outwards are frozen for the step, as in the generalized Cauchy point of Byrd,
Lu, Nocedal & Zhu 1995), projected backtracking line search with the Armijo condition, curvature-guarded history update.
Stopping tests are scipy L-BFGS-B's: max |projected gradient| <= gtol, or (f_k - f_{k+1}) / max(|f_k|, |f_{k+1}|, 1) <= ftol
(in two consecutive iterations, the second one a scaled steepest-descent step after the memory has been dropped: a
backtracking search, unlike scipy's Wolfe search, can return a crippled step from a poor quasi-Newton direction), or maxiter
iterations.  Problems that have stopped are dropped from the batch, so late iterations evaluate fewer points.
"""
import numpy as np


def _projected_gradient(x, g, lo, hi):
    pg = g.copy()
    pg[(x <= lo) & (g > 0)] = 0.0
    pg[(x >= hi) & (g < 0)] = 0.0
    return pg


def batched_lbfgsb(fun, X0, bounds=None, m=10, maxiter=1000, gtol=1e-5, ftol=1e7 * np.finfo(float).eps, maxls=20,
                   callback=None):
    """Minimise B independent problems in lock step.

    fun(X, idx) -> (f (b,), G (b, n)) evaluates the problems with indices ``idx`` (int array, b <= B) at the rows of X (b, n);
    non-finite values are treated as +inf (the step is shortened).  bounds: list of (lo, hi) per variable (shared by all
    problems; -inf / inf allowed, as for scipy) or None.
    Returns dict(x (B, n), fun (B,), nit (B,), nfev (total batched calls), status (B,) of
    'gtol' | 'ftol' | 'maxiter' | 'linesearch' | 'nonfinite-start')."""
    X = np.array(X0, dtype=np.float64, copy=True)
    if X.ndim == 1:
        X = X[None, :]
    B, n = X.shape
    lo = np.full(n, -np.inf) if bounds is None else np.array([-np.inf if b[0] is None else b[0] for b in bounds], dtype=np.float64)
    hi = np.full(n, np.inf) if bounds is None else np.array([np.inf if b[1] is None else b[1] for b in bounds], dtype=np.float64)
    X = np.clip(X, lo, hi)
    allidx = np.arange(B)
    F, G = fun(X, allidx)
    F = np.where(np.isfinite(F), F, np.inf).astype(np.float64)
    G = np.asarray(G, dtype=np.float64).copy()
    nfev = 1
    S = np.zeros((B, m, n))
    Yh = np.zeros((B, m, n))
    rho = np.zeros((B, m))
    nhist = np.zeros(B, dtype=int)
    nit = np.zeros(B, dtype=int)
    status = np.array(["running"] * B, dtype=object)
    status[~np.isfinite(F)] = "nonfinite-start"
    pg = np.stack([_projected_gradient(X[b], G[b], lo, hi) for b in range(B)])
    status[(status == "running") & (np.max(np.abs(pg), axis=1) <= gtol)] = "gtol"

    strikes = np.zeros(B, dtype=int)         # consecutive iterations whose decrease fell below ftol
    for _ in range(int(maxiter)):
        act = np.where(status == "running")[0]
        if act.size == 0:
            break
        # ---- search directions: two-loop recursion on the free variables, vectorised over the active problems (history is
        #      stored left-aligned: slots >= nhist[b] hold zeros and rho = 0, so they drop out of both loops)
        Xa, Ga = X[act], G[act]
        free = ~(((Xa <= lo) & (Ga > 0)) | ((Xa >= hi) & (Ga < 0)))
        q = np.where(free, Ga, 0.0)
        Sa, Ya, ra, ka = S[act], Yh[act], rho[act], nhist[act]
        alpha = np.zeros((act.size, m))
        for i in range(m - 1, -1, -1):
            alpha[:, i] = ra[:, i] * np.einsum("an,an->a", Sa[:, i], q)
            q = q - alpha[:, i, None] * Ya[:, i]
        newest = np.maximum(ka - 1, 0)
        ar = np.arange(act.size)
        sy = np.einsum("an,an->a", Sa[ar, newest], Ya[ar, newest])
        yy = np.einsum("an,an->a", Ya[ar, newest], Ya[ar, newest])
        gamma = np.where(ka > 0, sy / np.maximum(yy, 1e-300), 1.0 / np.maximum(np.linalg.norm(q, axis=1), 1e-300))
        r = gamma[:, None] * q
        for i in range(m):
            beta = ra[:, i] * np.einsum("an,an->a", Ya[:, i], r)
            r = r + Sa[:, i] * (alpha[:, i] - beta)[:, None]
        D = np.where(free, -r, 0.0)
        notdesc = ~(np.einsum("an,an->a", D, Ga) < 0.0)          # not a descent direction: restart from steepest descent
        if np.any(notdesc):
            D[notdesc] = np.where(free[notdesc], -Ga[notdesc], 0.0)
            bad = act[notdesc]
            nhist[bad] = 0
            S[bad], Yh[bad], rho[bad] = 0.0, 0.0, 0.0
        # ---- projected backtracking line search, all active problems together
        step = np.ones(act.size)
        done = np.zeros(act.size, dtype=bool)
        Xn, Fn, Gn = X[act].copy(), F[act].copy(), G[act].copy()
        for _ls in range(int(maxls)):
            todo = np.where(~done)[0]
            if todo.size == 0:
                break
            Xt = np.clip(X[act[todo]] + step[todo, None] * D[todo], lo, hi)
            ft, gt = fun(Xt, act[todo])
            nfev += 1
            ft = np.where(np.isfinite(ft), ft, np.inf)
            gt = np.asarray(gt, dtype=np.float64)
            # Armijo on the projected step: f(x+) <= f + c1 g.(x+ - x)
            dec = np.einsum("an,an->a", G[act[todo]], Xt - X[act[todo]])
            ok = np.isfinite(ft) & (ft <= F[act[todo]] + 1e-4 * dec) & np.all(np.isfinite(gt), axis=1)
            sel = todo[ok]
            Xn[sel], Fn[sel], Gn[sel] = Xt[ok], ft[ok], gt[ok]
            done[sel] = True
            step[todo[~ok]] *= 0.5
        # ---- accept / stop
        for a, b in enumerate(act):
            if not done[a]:
                status[b] = "linesearch"
                continue
            s, y = Xn[a] - X[b], Gn[a] - G[b]
            sy = np.dot(s, y)
            if sy > 1e-10 * np.dot(y, y) and sy > 0:
                if nhist[b] == m:
                    S[b, :-1], Yh[b, :-1], rho[b, :-1] = S[b, 1:].copy(), Yh[b, 1:].copy(), rho[b, 1:].copy()
                    nhist[b] -= 1
                k = nhist[b]
                S[b, k], Yh[b, k], rho[b, k] = s, y, 1.0 / sy
                nhist[b] += 1
            f_old = F[b]
            X[b], F[b], G[b] = Xn[a], Fn[a], Gn[a]
            nit[b] += 1
            pgb = _projected_gradient(X[b], G[b], lo, hi)
            if np.max(np.abs(pgb)) <= gtol:
                status[b] = "gtol"
            elif (f_old - F[b]) <= ftol * max(abs(f_old), abs(F[b]), 1.0):
                # scipy's L-BFGS-B can trust this test because its Wolfe line search never returns a crippled step; a
                # backtracking search can (a poor quasi-Newton direction, cut to a tiny step).  So the first strike only
                # drops the memory -- the next step is scaled steepest descent -- and a second consecutive one stops.
                strikes[b] += 1
                if strikes[b] >= 2:
                    status[b] = "ftol"
                else:
                    nhist[b] = 0
                    S[b], Yh[b], rho[b] = 0.0, 0.0, 0.0
            else:
                strikes[b] = 0
        if callback is not None:
            callback(X, F, status)
    status[status == "running"] = "maxiter"
    return {"x": X, "fun": F, "nit": nit, "nfev": nfev, "status": status}
