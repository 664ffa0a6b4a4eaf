"""TEST INFRASTRUCTURE ONLY -- deterministic synthetic inputs for the GPCSD hot path.

Geometries and hyperparameters of the BASELINE.json configs; LFP is a MODEL-MATCHED draw
``Y_r = Ls Z_r Lt^T + sqrt(sig2n) E_r`` (the regime where the reference's own solver-to-solver spread
is <= 1e-11).  Pure numpy; no reference import, so it also runs on the GPU box.
"""
import numpy as np

from . import gpcsd_oracle as O


def geometry_1d(nx=24, nt=50, ms_grid=False):
    """sim_from_gp_1D.py-style probe: 24 contacts over 2300 um."""
    x = np.linspace(0.0, 2300.0, nx)[:, None]
    t = (np.arange(nt, dtype=np.float64) if ms_grid else np.linspace(0.0, float(nt), nt))[:, None]
    return x, t


def geometry_neuropixels(nch=384, nt=250, dt=0.4):
    """neuropixels/extract_data.py:36-42 checkerboard: x=[16,48,0,32][ch%4], y=20*floor(ch/2)."""
    ch = np.arange(nch)
    xs = np.array([16.0, 48.0, 0.0, 32.0])[ch % 4]
    ys = 20.0 * np.floor(ch / 2)
    return np.stack([xs, ys], axis=1), (dt * np.arange(nt, dtype=np.float64))[:, None]


def geometry_grid_2d(nx1=4, nx2=12, nt=30):
    x1 = np.linspace(0.0, 48.0, nx1)
    x2 = np.linspace(0.0, 220.0, nx2)
    X = np.array([(a, b) for a in x1 for b in x2])
    return X, np.arange(nt, dtype=np.float64)[:, None]


def model_1d(x, t, a=None, b=None, ngl=100, sig2n=1e-2, unit_scale=True):
    """True parameters of simulation_studies/sim_from_gp_1D.py:41-47 (R=100, ell=200, SE(20, .5),
    Matern(5, .7)); sigma2_t divided by mean diag(Ks) so that the LFP has unit scale."""
    a = float(np.min(x)) if a is None else a
    b = float(np.max(x)) if b is None else b
    m = O.Model(1, O.Spatial1D(x, a, b, ngl), t, 100.0, (200.0,), [(O.KIND_SE, 20.0, 0.5), (O.KIND_MATERN, 5.0, 0.7)], sig2n)
    if unit_scale:
        tr = np.trace(m.Ks()) / len(x)
        m.temporal = [(k, e, s / tr) for k, e, s in m.temporal]
    return m


def model_2d(X, t, ngl1=20, ngl2=60, a1=None, b1=None, a2=None, b2=None, eps=1.0, sig2n=0.5,
             R=100.0, ell1=40.0, ell2=200.0, ell_se=5.0, ell_m=1.0, unit_scale=True):
    a1 = float(X[:, 0].min()) if a1 is None else a1
    b1 = float(X[:, 0].max()) if b1 is None else b1
    a2 = float(X[:, 1].min()) if a2 is None else a2
    b2 = float(X[:, 1].max()) if b2 is None else b2
    m = O.Model(2, O.Spatial2D(X, a1, b1, a2, b2, ngl1, ngl2), t, R, (ell1, ell2),
                [(O.KIND_SE, ell_se, 0.5), (O.KIND_MATERN, ell_m, 0.7)], sig2n, eps)
    if unit_scale:
        tr = np.trace(m.Ks()) / X.shape[0]
        m.temporal = [(k, e, s / tr) for k, e, s in m.temporal]
    return m


def matched_lfp(model, ntrials, seed):
    """Y_r = Ls Z_r Lt^T + sqrt(sig2n) E_r with Ls = Qs sqrt(max(ls,0)), Lt = Qt sqrt(max(lt,0))."""
    rng = np.random.default_rng(seed)
    ls, Qs = np.linalg.eigh(model.Ks(jitter=True))
    lt, Qt = np.linalg.eigh(model.Kt())
    Ls = Qs * np.sqrt(np.maximum(ls, 0.0))
    Lt = Qt * np.sqrt(np.maximum(lt, 0.0))
    nx, nt = Ls.shape[0], Lt.shape[0]
    Z = rng.standard_normal((nx, nt, ntrials))
    E = rng.standard_normal((nx, nt, ntrials))
    s = float(np.mean(np.atleast_1d(model.sig2n)))
    Y = np.einsum("ia,ajr->ijr", Ls, Z, optimize=True)
    Y = np.einsum("ijr,bj->ibr", Y, Lt, optimize=True)
    return np.ascontiguousarray(Y + np.sqrt(s) * E)


def perturbed(model, seed, scale=0.1):
    """theta = theta_true + scale*N(0,1) in log space (gradient evaluation point)."""
    rng = np.random.default_rng(seed)
    tp = O.pack_tparams(model)
    return O.unpack_tparams(model, tp + scale * rng.standard_normal(tp.shape))


def default_priors(model):
    """Flat prior-spec list in tparams order, following the reference's default constructors
    (gpcsd1d.py:50-62, gpcsd2d.py:64-79, covariances.py:40-48, 156-175, 243-255, 277-289)."""
    if model.dim == 1:
        xs = model.spatial.x.squeeze()
        dmin, span = np.min(np.diff(xs)), np.max(xs) - np.min(xs)
        pri = [("invgamma",) + O.invgamma_from_bounds(dmin, 0.5 * span),
               ("invgamma",) + O.invgamma_from_bounds(1.2 * dmin, 0.8 * span)]
        sd_n = 0.1
    else:
        X = model.spatial.x
        x1, x2 = np.sort(np.unique(X[:, 0])), np.sort(np.unique(X[:, 1]))
        sp = model.spatial
        mind = min(np.min(np.diff(x1)), np.min(np.diff(x2)))
        maxd = max(sp.b1 - sp.a1, sp.b2 - sp.a2)
        pri = [("invgamma",) + O.invgamma_from_bounds(mind, 0.5 * maxd),
               ("invgamma",) + O.invgamma_from_bounds(2.0 * np.min(np.diff(x1)), 2.0 * (x1.max() - x1.min())),
               ("invgamma",) + O.invgamma_from_bounds(2.0 * np.min(np.diff(x2)), (x2.max() - x2.min()))]
        sd_n = 1.0
    ts = np.asarray(model.t).flatten()
    tl, tu = 1.2 * np.min(np.diff(ts)), 0.8 * (ts.max() - ts.min())
    for _ in model.temporal:
        pri += [("invgamma",) + O.invgamma_from_bounds(tl, tu), ("halfnormal", 1.0)]
    pri += [("halfnormal", sd_n)] * len(np.atleast_1d(model.sig2n))
    return pri
