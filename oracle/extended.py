"""TEST INFRASTRUCTURE ONLY -- 80-bit (numpy longdouble) evaluation of the oracle's formulas.

The per-electrode-noise gradient contains (s_i - s_i') / (ls_i - ls_i') * Ns: for the near-null
spatial eigen-directions (ls ~ 1e-6) these core entries are ~1e5 times larger than the rest and cancel in the contraction
with dKs/dtheta, so ANY float64 evaluation of d loglik / d(R, ell) -- numpy's included -- carries ~1e-11 relative rounding
error even when both sides use bit-identical eigen-factors.  This module evaluates the same restatement
(oracle.gpcsd_oracle.loglik_and_grad, which follows gpcsd1d.py:113-128 and utility_functions.py:44-64) in extended
precision with the caller's float64 factors cast exactly, as the arbiter that says how much of a float64 difference is
conditioning and how much would be a defect.  Small shapes only (no BLAS for longdouble).
"""
import copy

import numpy as np

from . import gpcsd_oracle as O

LD = np.longdouble


def model_to_longdouble(om):
    sp = copy.copy(om.spatial)
    for k, v in vars(sp).items():
        if isinstance(v, np.ndarray) and v.dtype == np.float64:
            setattr(sp, k, v.astype(LD))
        elif isinstance(v, float):
            setattr(sp, k, LD(v))
    sig = np.asarray(om.sig2n).astype(LD) if np.ndim(om.sig2n) else LD(om.sig2n)
    return O.Model(om.dim, sp, np.asarray(om.t).astype(LD), LD(om.R), tuple(LD(e) for e in om.ells),
                   [(k, LD(e), LD(s)) for k, e, s in om.temporal], sig, LD(om.eps))


def loglik_and_grad_extended(om, lfp, factors):
    """(loglik, gradient) of the oracle's closed form evaluated in 80-bit arithmetic from float64 inputs and the given
    float64 factors (Qs, ls, Qt, lt); returned rounded to float64."""
    if np.finfo(LD).eps > 1e-18:
        raise RuntimeError("numpy longdouble is not extended precision on this platform")
    Qs, ls, Qt, lt = (np.asarray(f).astype(LD) for f in factors)
    nx = Qs.shape[0]
    first = [True]

    def eigh(K):
        # comp_eig_D calls eigh(Kt) first, then eigh(Ks) (utility_functions.py:58-59 order in the oracle restatement)
        if K.shape[0] == nx and (nx != Qt.shape[0] or not first[0]):
            return ls, Qs
        first[0] = False
        return lt, Qt
    ll, g = O.loglik_and_grad(model_to_longdouble(om), np.asarray(lfp).astype(LD), eigh=eigh)
    return float(ll), np.asarray(g, dtype=np.float64)
