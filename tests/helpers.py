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
