"""Drop-in API on the GPU (-m gpu): the reference-facing classes produce the reference's numbers."""
import os

import numpy as np
import pytest

from helpers import relerr
from test_gpu_engine import _model_from_golden_1d, _model_from_golden_2d, grad_tol

pytestmark = pytest.mark.gpu


def _api_model_1d(g, vec=False):
    from gpcsd_b200.covariances import GPCSDTemporalCovMatern, GPCSDTemporalCovSE
    from gpcsd_b200.gpcsd1d import GPCSD1D
    from gpcsd_b200.priors import GPCSDHalfNormalPrior
    np.random.seed(0)
    pri = [GPCSDHalfNormalPrior(0.1) for _ in range(24)] if vec else None
    m = GPCSD1D(g["lfp"], g["x"], g["t"], temporal_cov_list=[GPCSDTemporalCovSE(g["t"]), GPCSDTemporalCovMatern(g["t"])],
                sig2n_prior=pri)
    m.R['value'] = float(g["R"])
    m.spatial_cov.params['ell']['value'] = float(g["ell"])
    for tc, e, s in zip(m.temporal_cov_list, g["t_ell"], g["t_sigma2"]):
        tc.params['ell']['value'], tc.params['sigma2']['value'] = float(e), float(s)
    m.sig2n['value'] = np.array(g["sig2n"]) if vec else float(g["sig2n"])
    return m


def _api_model_2d(g):
    from gpcsd_b200.gpcsd2d import GPCSD2D
    np.random.seed(0)
    m = GPCSD2D(g["lfp"], g["x"], g["t"], ngl1=int(g["ngl1"]), ngl2=int(g["ngl2"]), eps=float(g["eps"]))
    m.R['value'] = float(g["R"])
    m.spatial_cov.params['ell1']['value'], m.spatial_cov.params['ell2']['value'] = float(g["ell1"]), float(g["ell2"])
    for tc, e, s in zip(m.temporal_cov_list, g["t_ell"], g["t_sigma2"]):
        tc.params['ell']['value'], tc.params['sigma2']['value'] = float(e), float(s)
    m.sig2n['value'] = float(g["sig2n"])
    return m


@pytest.mark.parametrize("name,vec", [("gpcsd1d_cfg1", False), ("gpcsd1d_vecnoise", True)])
def test_gpcsd1d_api_against_reference_golden(cuda_lib, golden_dir, name, vec):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    m = _api_model_1d(g, vec)
    ll = m.loglik()
    assert abs(float(ll) - float(g["loglik"])) < 1e-9 * abs(float(g["loglik"]))
    # covariance helper API (device kernels, host arrays out) vs reference matrices
    sc = m.spatial_cov
    assert relerr(sc.compKphi_1d(float(g["R"])), g["Ks"]) < 1e-12
    assert relerr(sc.compKphig_1d(g["z"], float(g["R"])), g["Kphig"]) < 1e-12
    assert relerr(sc.compKphi_1d(float(g["R"]), xp=g["z"]), g["Kphi_z"]) < 1e-12
    assert relerr(sc.compute_Ks(), g["Ks_csd"]) < 1e-14
    assert relerr(m.temporal_cov_list[0].compute_Kt(), g["Kt_se"]) < 1e-14
    assert relerr(m.temporal_cov_list[1].compute_Kt(), g["Kt_matern"]) < 1e-14
    z, tt = g["z"], g["t"]
    m.predict(z, tt, type="both")
    npred = g["csd_pred"].shape[2]
    assert m.csd_pred.shape == (22, g["t"].shape[0], g["lfp"].shape[2]) and len(m.csd_pred_list) == 2
    assert m.x_pred is z and m.t_pred is tt
    for key in ("csd_pred", "lfp_pred"):
        assert relerr(getattr(m, key)[:, :, :npred], g[key]) < 1e-6      # vs the reference's dense inverse
    m.predict(g["z"], g["t"])                                             # default type="csd"
    assert relerr(m.csd_pred[:, :, :npred], g["csd_pred"]) < 1e-6


def test_gpcsd2d_api_against_reference_golden(cuda_lib, golden_dir):
    g = np.load(os.path.join(golden_dir, "gpcsd2d_small.npz"))
    m = _api_model_2d(g)
    assert abs(float(m.loglik()) - float(g["loglik"])) < 1e-9 * abs(float(g["loglik"]))
    sc = m.spatial_cov
    assert relerr(sc.compKphi_2d(float(g["R"]), float(g["eps"])), g["Ks"]) < 1e-12
    assert relerr(sc.compKphig_2d(g["z"], float(g["R"]), float(g["eps"])), g["Kphig"]) < 1e-12
    assert relerr(sc.compKphi_2d(float(g["R"]), float(g["eps"]), xp=g["z"]), g["Kphi_z"]) < 1e-12
    assert relerr(sc.compute_Ks(), g["Ks_csd"]) < 1e-14
    m.predict(g["z"], g["t"], type="both")
    for key in ("csd_pred", "lfp_pred"):
        assert relerr(getattr(m, key), g[key]) < 1e-6
        assert relerr(getattr(m, key + "_list")[0], g[key + "_0"]) < 1e-6


def test_obj_fun_and_grad_matches_oracle_objective(cuda_lib, golden_dir):
    """nll and its log-space gradient (priors + exp transforms, gpcsd1d.py:153-191) vs the oracle's."""
    from oracle import gpcsd_oracle as O, synth
    g = np.load(os.path.join(golden_dir, "gpcsd1d_cfg1.npz"))
    m = _api_model_1d(g)
    om = _model_from_golden_1d(g)
    pri = synth.default_priors(om)
    tp = O.pack_tparams(om) + 0.05
    f, gr = m.obj_fun_and_grad(tp)
    f_o, gr_o = O.obj_and_grad(om, g["lfp"], tp, pri)
    assert abs(f - f_o) < 1e-9 * abs(f_o)
    assert np.max(np.abs(gr - gr_o) / np.maximum(np.abs(gr_o), 1e-9 * np.max(np.abs(gr_o)))) < 1e-8
    assert abs(m.obj_fun(tp) - f_o) < 1e-9 * abs(f_o)
    f2, gr2 = m.obj_fun_and_grad(tp, fix_R=True)
    assert gr2[0] == 0.0


def test_fit_recovers_and_improves(cuda_lib):
    """fit(): multi-start bounded L-BFGS-B driven by the fused CUDA objective+gradient; the MAP objective
    at the fitted parameters must be no worse than at the data-generating parameters."""
    from gpcsd_b200.gpcsd1d import GPCSD1D
    from oracle import gpcsd_oracle as O, synth
    x, t = synth.geometry_1d(24, 40)
    om = synth.model_1d(x, t, sig2n=1e-2)
    lfp = synth.matched_lfp(om, 30, 11)
    np.random.seed(2)
    m = GPCSD1D(lfp, x, t)
    tp_true = O.pack_tparams(om)
    f_true = m.obj_fun(tp_true)
    m.fit(n_restarts=2, options={'maxiter': 60, 'disp': False, 'gtol': 1e-5, 'ftol': 1e7 * np.finfo(float).eps})
    tp_fit = np.log(np.array([m.R['value'] / 100, m.spatial_cov.params['ell']['value'] / 100] +
                             [v for tc in m.temporal_cov_list for v in (tc.params['ell']['value'], tc.params['sigma2']['value'])] +
                             [m.sig2n['value']]))
    f_fit = m.obj_fun(tp_fit)
    assert np.isfinite(f_fit) and f_fit <= f_true + 1e-6 * abs(f_true)
    assert 0.3 * 1e-2 < m.sig2n['value'] < 3e-2            # noise variance is well identified


def test_sample_prior_and_comp_eig_D(cuda_lib):
    from gpcsd_b200.gpcsd1d import GPCSD1D
    from gpcsd_b200.utility_functions import comp_eig_D
    from oracle import gpcsd_oracle as O
    np.random.seed(3)
    x = np.linspace(0, 2300, 24)[:, None]
    t = np.linspace(0, 30, 31)[:, None]
    m = GPCSD1D(np.zeros((24, 31, 1)), x, t)
    m.spatial_cov.params['ell']['value'] = 300.0
    np.random.seed(7)
    csd = m.sample_prior(5)
    assert csd.shape == (24, 31, 5) and np.all(np.isfinite(csd))
    # same draws through the reference's formula on the host
    Kt = sum(O.compute_Kt(tc.KIND, tc.params['ell']['value'], tc.params['sigma2']['value'], t) for tc in m.temporal_cov_list)
    Lt = np.linalg.cholesky(Kt)
    Ls = np.linalg.cholesky(O.compute_Ks_1d(x, 300.0) + 1e-8 * np.eye(24))
    np.random.seed(7)
    ref = np.stack([Ls @ np.random.normal(0, 1, (24, 31)) @ Lt.T for _ in range(5)], axis=2)
    # cond(Ks_csd + 1e-8 I) ~ 1e8: round-off level differences in Ks are amplified by the Cholesky factor
    assert relerr(csd, ref) < 1e-7
    rng = np.random.default_rng(0)
    Ks = rng.standard_normal((9, 9)); Ks = Ks @ Ks.T
    Kt2 = rng.standard_normal((13, 13)); Kt2 = Kt2 @ Kt2.T
    sv = rng.uniform(0.1, 0.4, 9)
    for sig in (0.3, sv):
        Qs, Qt, D = comp_eig_D(Ks, Kt2, sig)
        _, _, D_o, ls, lt = O.comp_eig_D(Ks, Kt2, sig)
        assert relerr(D, D_o) < 1e-12 and D.shape == (9 * 13,)
        assert relerr((Qs * ls) @ Qs.T, Ks) < 1e-12 and relerr((Qt * lt) @ Qt.T, Kt2) < 1e-12


def test_update_lfp_reuploads(cuda_lib, golden_dir):
    g = np.load(os.path.join(golden_dir, "gpcsd1d_lownoise.npz"))
    m = _api_model_1d(g)
    ll1 = float(m.loglik())
    m.update_lfp(2.0 * g["lfp"], g["t"])
    ll2 = float(m.loglik())
    assert abs(ll1 - float(g["loglik"])) < 1e-9 * abs(ll1) and ll2 != ll1
    m.update_lfp(g["lfp"], g["t"])
    assert abs(float(m.loglik()) - ll1) < 1e-12 * abs(ll1)


def test_in_place_edit_of_lfp_is_seen(cuda_lib, golden_dir):
    """The reference re-reads self.lfp on every call; an in-place edit of the same array object must not leave a stale device
    copy (content fingerprint in GPCSDModelBase._get_engine), and invalidate() forces a re-upload unconditionally."""
    g = np.load(os.path.join(golden_dir, "gpcsd1d_lownoise.npz"))
    m = _api_model_1d(g)
    ll1 = float(m.loglik())
    m.lfp *= 2.0                                       # same object, new contents
    ll2 = float(m.loglik())
    assert ll2 != ll1
    m.lfp *= 0.5
    m.invalidate()
    assert abs(float(m.loglik()) - ll1) < 1e-12 * abs(ll1)


def test_fit_concurrent_workers_match_sequential(cuda_lib):
    """fit(n_workers=2) -- restarts on two threads/streams/engines sharing the uploaded LFP -- must end at exactly the
    parameters of the sequential fit (same starts, deterministic kernels)."""
    from gpcsd_b200.gpcsd1d import GPCSD1D
    from oracle import synth
    x, t = synth.geometry_1d(24, 40)
    om = synth.model_1d(x, t, sig2n=1e-2)
    lfp = synth.matched_lfp(om, 25, 17)
    opts = {'maxiter': 40, 'disp': False, 'gtol': 1e-5, 'ftol': 1e7 * np.finfo(float).eps}
    out = []
    for workers in (1, 2):
        np.random.seed(5)
        m = GPCSD1D(lfp, x, t)
        m.fit(n_restarts=4, options=opts, n_workers=workers, lockstep=False)
        p = m.extract_model_params()
        out.append(np.array([p['R'], p['spatial_ell'], p['sig2n']] + p['temporal_ell_list'] + p['temporal_sigma2_list']))
    assert np.array_equal(out[0], out[1])


def test_fit_fix_R_and_changing_trial_counts(cuda_lib):
    """fix_R=True keeps R (gpcsd1d.py:161,195-196,234); update_lfp with a different trial count re-sizes the device
    buffers; both give oracle-consistent likelihoods afterwards."""
    from gpcsd_b200.gpcsd1d import GPCSD1D
    from oracle import gpcsd_oracle as O, synth
    x, t = synth.geometry_1d(24, 36)
    om = synth.model_1d(x, t, sig2n=1e-2)
    lfp = synth.matched_lfp(om, 21, 8)
    np.random.seed(9)
    m = GPCSD1D(lfp[:, :, :12], x, t)
    m.R['value'] = 123.0
    m.fit(n_restarts=2, fix_R=True, options={'maxiter': 25, 'disp': False, 'gtol': 1e-5, 'ftol': 1e7 * np.finfo(float).eps})
    assert m.R['value'] == 123.0
    for n in (21, 5, 16):                                   # grow, shrink, grow: buffers follow the trial count
        m.update_lfp(lfp[:, :, :n], t)
        p = m.extract_model_params()
        om_fit = O.Model(1, om.spatial, t, p['R'], (p['spatial_ell'],),
                         [(0, p['temporal_ell_list'][0], p['temporal_sigma2_list'][0]), (1, p['temporal_ell_list'][1], p['temporal_sigma2_list'][1])],
                         float(p['sig2n']))
        ll, ll_o = float(m.loglik()), O.loglik(om_fit, lfp[:, :, :n])
        assert abs(ll - ll_o) < 1e-9 * abs(ll_o)
        m.predict(x, t)
        assert m.csd_pred.shape == (24, 36, n)


def test_lockstep_fit_matches_oracle_driven_scipy_fit(cuda_lib):
    """fit() with all restarts in lock step (restart-batched native evaluations) against the reference procedure restated on
    the CPU: scipy L-BFGS-B per restart (gpcsd1d.py:193-211) on the ORACLE objective/gradient, same prior-sampled starts,
    same bounds and options.  Both stop on the same gtol / ftol tests, so the optima agree to that tolerance."""
    import scipy.optimize
    from gpcsd_b200.gpcsd1d import GPCSD1D
    from oracle import gpcsd_oracle as O, synth
    x, t = synth.geometry_1d(24, 40)
    om = synth.model_1d(x, t, sig2n=1e-2)
    lfp = synth.matched_lfp(om, 30, 11)
    opts = {'maxiter': 300, 'disp': False, 'gtol': 1e-5, 'ftol': 1e7 * np.finfo(float).eps}
    np.random.seed(21)
    m = GPCSD1D(lfp, x, t)
    state = np.random.get_state()
    m.fit(n_restarts=6, options=opts)                                  # lock step (default for L-BFGS-B)
    info = m._last_fit_info
    tp_fit = np.log(np.array([m.R['value'] / 100, m.spatial_cov.params['ell']['value'] / 100] +
                             [v for tc in m.temporal_cov_list for v in (tc.params['ell']['value'], tc.params['sigma2']['value'])] +
                             [m.sig2n['value']]))
    f_lock = m.obj_fun(tp_fit)
    # the reference procedure on the oracle, from the same starts (same RNG state -> same prior draws)
    np.random.set_state(state)
    starts = [m._sample_tparams0(False) for _ in range(6)]
    pri = synth.default_priors(om)
    bounds = m._bounds()
    best = np.inf
    for s0 in starts:
        r = scipy.optimize.minimize(lambda tp: O.obj_and_grad(om, lfp, tp, pri), s0, jac=True, method="L-BFGS-B", bounds=bounds,
                                    options={k: v for k, v in opts.items() if k != 'disp'})
        if np.isfinite(r.fun):
            best = min(best, float(r.fun))
    print("\n[lock-step fit] nll %.6f vs oracle-driven scipy %.6f; %d batched calls for 6 restarts (iterations per restart %s)"
          % (f_lock, best, info["batched_calls"], list(info["nit"])))
    assert abs(f_lock - best) <= 1e-5 * abs(best)
    # and against the per-restart scipy path of this package (lockstep=False) from the same starts
    np.random.seed(21)
    m2 = GPCSD1D(lfp, x, t)
    m2.fit(n_restarts=6, options=opts, lockstep=False, n_workers=1)
    tp2 = np.log(np.array([m2.R['value'] / 100, m2.spatial_cov.params['ell']['value'] / 100] +
                          [v for tc in m2.temporal_cov_list for v in (tc.params['ell']['value'], tc.params['sigma2']['value'])] +
                          [m2.sig2n['value']]))
    assert abs(m2.obj_fun(tp2) - f_lock) <= 1e-5 * abs(f_lock)


def test_lockstep_fit_per_electrode_noise_2d_and_failures(cuda_lib):
    """Lock-step fit with per-electrode noise (P = 30) and for the 2-D model; a start with non-finite objective is dropped
    like a restart that raises in the reference (gpcsd1d.py:219)."""
    from gpcsd_b200.gpcsd1d import GPCSD1D
    from gpcsd_b200.gpcsd2d import GPCSD2D
    from gpcsd_b200.priors import GPCSDHalfNormalPrior
    from oracle import synth
    x, t = synth.geometry_1d(24, 40)
    rng = np.random.default_rng(0)
    om = synth.model_1d(x, t, sig2n=1e-2 * np.exp(0.3 * rng.standard_normal(24)))
    lfp = synth.matched_lfp(om, 40, 3)
    np.random.seed(4)
    m = GPCSD1D(lfp, x, t, sig2n_prior=[GPCSDHalfNormalPrior(0.1) for _ in range(24)])
    f0 = m.obj_fun(m._sample_tparams0(False))
    m.fit(n_restarts=3, options={'maxiter': 40, 'disp': False, 'gtol': 1e-5, 'ftol': 1e7 * np.finfo(float).eps})
    assert len(m.sig2n['value']) == 24 and np.all(np.isfinite(m.sig2n['value']))
    assert m.obj_fun(np.log(np.array([m.R['value'] / 100, m.spatial_cov.params['ell']['value'] / 100] +
                                     [v for tc in m.temporal_cov_list for v in (tc.params['ell']['value'], tc.params['sigma2']['value'])] +
                                     list(m.sig2n['value'])))) < f0
    X, t2 = synth.geometry_grid_2d(3, 6, 16)
    om2 = synth.model_2d(X, t2, ngl1=6, ngl2=10, sig2n=0.3)
    lfp2 = synth.matched_lfp(om2, 10, 5)
    np.random.seed(6)
    m2 = GPCSD2D(lfp2, X, t2, ngl1=6, ngl2=10)
    m2.fit(n_restarts=2, options={'maxiter': 15, 'disp': False, 'gtol': 1e-5, 'ftol': 1e7 * np.finfo(float).eps})
    assert np.isfinite(float(m2.loglik()))
