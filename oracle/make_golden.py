"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (needs /root/reference):  python -m oracle.make_golden
Each fixture stores explicit inputs (arrays + hyperparameters, never RNG replay -- SURVEY.md 9.8) and
the outputs the reference produced for them through its own public API (GPCSD1D/GPCSD2D objects,
covariance classes, helper functions).  Gradients are 4th-order central finite differences of the
reference's own ``loglik`` (the reference's autograd gradient cannot run here).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import synth  # noqa: E402
from oracle import gpcsd_oracle as O  # noqa: E402
from oracle.ref_shim import import_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def _set_1d(m, om):
    m.R['value'] = om.R
    m.spatial_cov.params['ell']['value'] = om.ells[0]
    for tc, (_, ell, s2) in zip(m.temporal_cov_list, om.temporal):
        tc.params['ell']['value'] = ell
        tc.params['sigma2']['value'] = s2
    m.sig2n['value'] = om.sig2n


def _set_2d(m, om):
    m.R['value'] = om.R
    m.spatial_cov.params['ell1']['value'] = om.ells[0]
    m.spatial_cov.params['ell2']['value'] = om.ells[1]
    for tc, (_, ell, s2) in zip(m.temporal_cov_list, om.temporal):
        tc.params['ell']['value'] = ell
        tc.params['sigma2']['value'] = s2
    m.sig2n['value'] = om.sig2n
    m.eps = om.eps


def _natural(om):
    v = [om.R] + list(om.ells)
    for _, e, s in om.temporal:
        v += [e, s]
    return np.array(v + list(np.atleast_1d(om.sig2n)), dtype=np.float64)


def _from_natural(om, v):
    ns = len(om.ells)
    p = 1 + ns
    temporal = []
    for k, _, _ in om.temporal:
        temporal.append((k, v[p], v[p + 1]))
        p += 2
    sig = v[p] if np.ndim(om.sig2n) == 0 else np.array(v[p:])
    return O.Model(om.dim, om.spatial, om.t, v[0], tuple(v[1:1 + ns]), temporal, sig, om.eps)


def _fd_grad(ref_model, setter, om, rel_h=1e-4):
    v0 = _natural(om)
    g = np.zeros_like(v0)

    def f(v):
        setter(ref_model, _from_natural(om, v))
        return float(ref_model.loglik())

    for k in range(len(v0)):
        h = rel_h * v0[k]
        e = np.zeros_like(v0)
        e[k] = h
        g[k] = (-f(v0 + 2 * e) + 8 * f(v0 + e) - 8 * f(v0 - e) + f(v0 - 2 * e)) / (12 * h)
    setter(ref_model, om)
    return g


def _temporal_arrays(om):
    return (np.array([k for k, _, _ in om.temporal]), np.array([e for _, e, _ in om.temporal]),
            np.array([s for _, _, s in om.temporal]))


def golden_1d(g, name, nt, N, sig2n, npred, seed, vec=False):
    x, t = synth.geometry_1d(24, nt)
    om = synth.model_1d(x, t, sig2n=sig2n)
    if vec:
        rng = np.random.default_rng(seed + 99)
        om.sig2n = sig2n * np.exp(0.3 * rng.standard_normal(24))
    lfp = synth.matched_lfp(om, N, seed)
    np.random.seed(0)
    tcl = [g.covariances.GPCSDTemporalCovSE(t), g.covariances.GPCSDTemporalCovMatern(t)]
    pri = [g.priors.GPCSDHalfNormalPrior(0.1) for _ in range(24)] if vec else None
    m = g.gpcsd1d.GPCSD1D(lfp, x, t, temporal_cov_list=tcl, sig2n_prior=pri)
    _set_1d(m, om)
    ll = float(m.loglik())
    z = np.linspace(100.0, 2200.0, 22)[:, None]
    m.predict(z, t, type="both")
    kinds, tells, ts2 = _temporal_arrays(om)
    out = dict(x=x, t=t, lfp=lfp, a=m.a, b=m.b, ngl=m.ngl, R=om.R, ell=om.ells[0], t_kind=kinds, t_ell=tells,
               t_sigma2=ts2, sig2n=np.asarray(om.sig2n), loglik=ll, z=z,
               Ks=m.spatial_cov.compKphi_1d(om.R), Kphig=m.spatial_cov.compKphig_1d(z, om.R),
               Kphi_z=m.spatial_cov.compKphi_1d(om.R, xp=z), Ks_csd=m.spatial_cov.compute_Ks(),
               Kt_se=m.temporal_cov_list[0].compute_Kt(), Kt_matern=m.temporal_cov_list[1].compute_Kt(),
               csd_pred=m.csd_pred[:, :, :npred], lfp_pred=m.lfp_pred[:, :, :npred],
               csd_pred_0=m.csd_pred_list[0][:, :, :npred], csd_pred_1=m.csd_pred_list[1][:, :, :npred],
               lfp_pred_0=m.lfp_pred_list[0][:, :, :npred], lfp_pred_1=m.lfp_pred_list[1][:, :, :npred],
               gl_x=m.spatial_cov.gl_x, gl_w=m.spatial_cov.gl_w)
    if not vec:
        out["grad_fd_natural"] = _fd_grad(m, _set_1d, om)
        # gradient evaluation point theta_true + 0.1 N(0,1)
        om2 = synth.perturbed(om, seed + 1)
        _set_1d(m, om2)
        out["pert_natural"] = _natural(om2)
        out["pert_loglik"] = float(m.loglik())
        out["pert_grad_fd_natural"] = _fd_grad(m, _set_1d, om2)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "loglik", ll)


def golden_2d(g, name, seed):
    X, t = synth.geometry_grid_2d(4, 12, 30)
    om = synth.model_2d(X, t, ngl1=10, ngl2=30, eps=20.0, sig2n=0.3, R=60.0, ell1=30.0, ell2=70.0, ell_se=8.0, ell_m=2.0)
    lfp = synth.matched_lfp(om, 6, seed)
    np.random.seed(0)
    m = g.gpcsd2d.GPCSD2D(lfp, X, t, ngl1=10, ngl2=30, eps=20.0)
    _set_2d(m, om)
    ll = float(m.loglik())
    z = X[::5] + np.array([3.0, 7.0])
    m.predict(z, t, type="both")
    kinds, tells, ts2 = _temporal_arrays(om)
    sc = m.spatial_cov
    out = dict(x=X, t=t, lfp=lfp, a1=m.a1, b1=m.b1, a2=m.a2, b2=m.b2, ngl1=10, ngl2=30, eps=m.eps, R=om.R,
               ell1=om.ells[0], ell2=om.ells[1], t_kind=kinds, t_ell=tells, t_sigma2=ts2, sig2n=om.sig2n,
               loglik=ll, z=z, Ks=sc.compKphi_2d(om.R, m.eps), Kphig=sc.compKphig_2d(z, om.R, m.eps),
               Kphi_z=sc.compKphi_2d(om.R, m.eps, xp=z), Ks_csd=sc.compute_Ks(),
               gl_x_grid=sc.gl_x_grid, gl_w_prod=sc.gl_w_prod,
               csd_pred=m.csd_pred, lfp_pred=m.lfp_pred, csd_pred_0=m.csd_pred_list[0], csd_pred_1=m.csd_pred_list[1],
               lfp_pred_0=m.lfp_pred_list[0], lfp_pred_1=m.lfp_pred_list[1],
               grad_fd_natural=_fd_grad(m, _set_2d, om))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "loglik", ll)


def golden_helpers(g):
    """Helper-function vectors: forward models, grids, mykron, comp_eig_D, priors, tCSD."""
    rng = np.random.default_rng(5)
    uf, fm, pr, pc = g.utility_functions, g.forward_models, g.priors, g.predict_csd
    r = rng.standard_normal((5, 7)) * 300
    A = rng.standard_normal((3, 4))
    B = rng.standard_normal((2, 5))
    Ks = rng.standard_normal((6, 6)); Ks = Ks @ Ks.T
    Kt = rng.standard_normal((5, 5)); Kt = Kt @ Kt.T
    sv = rng.uniform(0.1, 0.5, 6)
    Qs, Qt, Dv = uf.comp_eig_D(Ks, Kt, 0.3)
    Qs2, Qt2, Dv2 = uf.comp_eig_D(Ks, Kt, sv)
    xd = np.linspace(0, 2300, 40)[:, None]
    zz = np.linspace(0, 2300, 9)[:, None]
    csd = rng.standard_normal((40, 6))
    x1 = np.linspace(0, 40, 5)[:, None]; x2 = np.linspace(0, 100, 8)[:, None]
    arr2 = rng.standard_normal((5, 8, 3))
    z2 = rng.uniform(0, 40, (4, 2))
    lf = rng.standard_normal((6, 5, 2))
    lf4 = rng.standard_normal((3, 6, 5, 2))
    ig = pr.GPCSDInvGammaPrior(); ig.set_params(3.0, 40.0)
    hn = pr.GPCSDHalfNormalPrior(0.7)
    xs = np.array([0.3, 2.0, 11.0])
    grid = uf.expand_grid(x1, x2)
    perm = rng.permutation(grid.shape[0])
    out = dict(r=r, b1d=fm.b_fwd_1d(r, 80.0), b2d=fm.b_fwd_2d(r, r.T[:5, :7] if False else r * 0.5, 80.0, 20.0),
               A=A, B=B, kron=uf.mykron(A, B), Ks=Ks, Kt=Kt, sv=sv, Dvec=Dv, Dvec_vec=Dv2,
               xd=xd, zz=zz, csd=csd, fwd1d=fm.fwd_model_1d(csd, xd, zz, 120.0, varsigma=0.4),
               x1=x1, x2=x2, arr2=arr2, z2=z2, fwd2d=fm.fwd_model_2d(arr2, x1, x2, z2, 60.0, 10.0),
               lf=lf, tcsd1=pc.predictcsd_trad_1d(lf), lf4=lf4, tcsd2=pc.predictcsd_trad_2d(lf4),
               ig_alpha=ig.alpha, ig_beta=ig.beta, xs=xs, ig_lpdf=np.array([ig.lpdf(v) for v in xs]),
               hn_lpdf=np.array([hn.lpdf(v) for v in xs]), grid=grid, grid_perm=grid[perm],
               grid_sorted=uf.sort_grid(grid[perm]), norm_in=lf, norm_out=uf.normalize(lf))
    np.savez_compressed(os.path.join(OUT, "helpers.npz"), **out)
    print("helpers ok")


def main():
    g = import_reference()
    os.makedirs(OUT, exist_ok=True)
    golden_1d(g, "gpcsd1d_cfg1", nt=50, N=50, sig2n=1e-2, npred=8, seed=1000)
    golden_1d(g, "gpcsd1d_lownoise", nt=40, N=9, sig2n=1e-4, npred=9, seed=1001)
    golden_1d(g, "gpcsd1d_vecnoise", nt=40, N=11, sig2n=1e-2, npred=11, seed=1002, vec=True)
    golden_2d(g, "gpcsd2d_small", seed=3000)
    golden_helpers(g)


if __name__ == "__main__":
    main()
