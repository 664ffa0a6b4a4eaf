"""TEST INFRASTRUCTURE ONLY -- multiprecision arbiter for the posterior mean (gpcsd1d.py:248-293), scalar noise.

The reference's predict multiplies out a dense (nx nt)^2 inverse (gpcsd1d.py:263-265), which loses about log10 cond(K) digits;
the engine and oracle.predict_kron use the algebraically identical Kronecker form.  When the two float64 results differ at
1e-6 (cond ~ 1e9), this module says which one is right: K = Ks (x) Kt + sig2n I is assembled from the SAME float64 covariance
matrices in 40-digit arithmetic (mpmath), K^-1 Y is solved there, and only the final, well-conditioned contraction with the
cross-covariances runs in extended precision.  Small shapes only (mpmath LU of an (nx nt) x (nx nt) matrix)."""
import numpy as np

from . import gpcsd_oracle as O


def predict_arbiter(model, lfp, z, kind="csd", dps=40):
    import mpmath as mp
    if np.ndim(model.sig2n):
        raise ValueError("scalar noise only: with per-electrode noise the reference's D is not Ks (x) Kt + diag")
    mp.mp.dps = dps
    lfp = np.atleast_3d(lfp)
    nx, nt, N = lfp.shape
    Ks, Kt, s = model.Ks(jitter=False), model.Kt(), float(model.sig2n)        # no jitter in predict (gpcsd1d.py:258)
    n = nx * nt
    K = mp.matrix(n, n)
    for i in range(nx):
        for a in range(nx):
            ksa = mp.mpf(float(Ks[i, a]))
            for j in range(nt):
                for b in range(nt):
                    K[i * nt + j, a * nt + b] = ksa * mp.mpf(float(Kt[j, b]))
    for d in range(n):
        K[d, d] += mp.mpf(s)
    W = np.zeros((nx, nt, N), dtype=np.longdouble)
    for r in range(N):
        w = mp.lu_solve(K, mp.matrix([mp.mpf(float(v)) for v in lfp[:, :, r].reshape(-1)]))
        W[:, :, r] = np.array([np.longdouble(mp.nstr(w[q], 25)) for q in range(n)]).reshape(nx, nt)
    Kc = (model.Kphig(z) if kind == "csd" else model.Ks(xp=z)).astype(np.longdouble)            # (nx, nz)
    parts = []
    for knd, ell, s2 in model.temporal:
        Kts = O.compute_Kt(knd, ell, s2, model.t, model.t).astype(np.longdouble)                 # (nt*, nt), t* = t
        # mykron(Kc, Kts).T @ invy  ->  out[z, j*, r] = sum_{i,j} Kc[i, z] Kts[j, j*] W[i, j, r]   (gpcsd1d.py:279)
        parts.append(np.einsum("iz,jk,ijr->zkr", Kc, Kts, W).astype(np.float64))
    return parts
