"""Shared helpers for the parity tests (oracle model <-> engine inputs)."""
import numpy as np


def engine_from_oracle(om, lfp=None, group=None):
    from gpcsd_b200.engine import HyperParams, KronEngine
    sp = om.spatial
    if om.dim == 1:
        quad = dict(gl_x=sp.gl_x, gl_w=sp.gl_w)
    else:
        quad = dict(gl_x1=sp.gl_x1, gl_w1=sp.gl_w1, gl_x2=sp.gl_x2, gl_w2=sp.gl_w2)
    eng = KronEngine(om.dim, sp.x, om.t, quad, group=group)
    if lfp is not None:
        eng.set_lfp(lfp)
    return eng, hp_from_oracle(om)


def hp_from_oracle(om):
    from gpcsd_b200.engine import HyperParams
    return HyperParams(R=om.R, ells=tuple(om.ells), temporal=list(om.temporal), sig2n=om.sig2n, eps=om.eps)


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def scaled_rel(g, go):
    """Component-wise relative error with a floor of 1e-6 of the largest component (components that vanish by symmetry)."""
    g, go = np.asarray(g, dtype=np.float64), np.asarray(go, dtype=np.float64)
    return np.abs(g - go) / np.maximum(np.abs(go), 1e-6 * np.max(np.abs(go)))


def ulp_sensitivity(om, lfp, nseeds=4):
    """How much the REFERENCE FORMULA's own gradient moves when the matrices handed to np.linalg.eigh are perturbed by at
    most ONE ulp per entry (symmetric, random signs) -- a perturbation below the backward error of any eigensolver, LAPACK
    included.  With per-electrode noise the function depends on eigenvector identity (utility_functions.py:54-57), and this
    is the floor below which two correct implementations cannot be expected to agree: 1.2e-8 at the configs[1] shape,
    2.5e-8 on the golden per-electrode case (LAPACK's three drivers share dsytrd, so their mutual spread -- 1e-10 / 2e-9 --
    understates it).  Returns the max scaled relative change over `nseeds` perturbations."""
    from oracle import gpcsd_oracle as O
    g0 = O.loglik_and_grad(om, lfp)[1]
    worst = 0.0
    for seed in range(nseeds):
        rng = np.random.default_rng(seed)

        def eigh(K):
            E = rng.uniform(-1.0, 1.0, K.shape)
            return np.linalg.eigh(K + np.finfo(np.float64).eps * np.abs(K) * 0.5 * (E + E.T))
        worst = max(worst, float(np.max(scaled_rel(O.loglik_and_grad(om, lfp, eigh=eigh)[1], g0))))
    return worst
