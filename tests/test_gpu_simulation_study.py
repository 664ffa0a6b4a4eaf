"""The reference's only "tests" are its simulation scripts.  This follows the flow of
simulation_studies/sim_from_gp_1D.py:25-110 through the DROP-IN package name (`from gpcsd.gpcsd1d import ...`,
star imports included) on the GPU: sample CSD from the GP prior on a dense grid, push it through the forward
model, add noise, normalise, predict at the true hyperparameters (and after a short fit), and score the
recovered CSD against the truth and against the traditional second-difference estimator."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _r2(pred, true):
    return 1.0 - np.sum((pred - true) ** 2) / np.sum((true - true.mean()) ** 2)


def test_sim_from_gp_1d_flow_through_dropin_names(cuda_lib):
    sys.path.insert(0, os.path.join(ROOT, "dropin"))
    import scipy.interpolate
    from gpcsd.gpcsd1d import GPCSD1D                      # noqa: E402  (alias of gpcsd_b200.gpcsd1d)
    ns = {}
    exec("from gpcsd.covariances import *", ns)            # scripts rely on this star import's re-exports
    GPCSDTemporalCovSE, GPCSDTemporalCovMatern = ns["GPCSDTemporalCovSE"], ns["GPCSDTemporalCovMatern"]
    fwd_model_1d = ns["fwd_model_1d"]                      # arrives via covariances' own `from ...forward_models import *`
    assert "GPCSDInvGammaPrior" in ns and "b_fwd_2d" in ns and "expand_grid" in ns and "np" in ns
    from gpcsd.predict_csd import predictcsd_trad_1d       # noqa: E402
    from gpcsd.utility_functions import normalize          # noqa: E402
    np.random.seed(1)
    ntrials, a, b, nt, nx, nz = 20, 0, 2300, 60, 24, 100
    t = np.linspace(0, nt, nt)[:, None]
    x = np.linspace(a, b, nx)[:, None]
    xshort = x[1:-1]
    z = np.linspace(a, b, nz)[:, None]
    true = dict(R=100, ell=200, se=(20.0, 0.5), mat=(5.0, 0.7), sig2n=0.0001)

    def set_true(m):
        m.R['value'] = true["R"]
        m.sig2n['value'] = true["sig2n"]
        m.spatial_cov.params['ell']['value'] = true["ell"]
        m.temporal_cov_list[0].params['ell']['value'], m.temporal_cov_list[0].params['sigma2']['value'] = true["se"]
        m.temporal_cov_list[1].params['ell']['value'], m.temporal_cov_list[1].params['sigma2']['value'] = true["mat"]

    gen = GPCSD1D(np.zeros((nz, nt)), z, t, temporal_cov_list=[GPCSDTemporalCovSE(t), GPCSDTemporalCovMatern(t)])
    set_true(gen)
    csd = gen.sample_prior(2 * ntrials)
    assert csd.shape == (nz, nt, 2 * ntrials)
    csd_interior = np.zeros((nx - 2, nt, 2 * ntrials))
    lfp = np.zeros((nx, nt, 2 * ntrials))
    for trial in range(2 * ntrials):
        csd_interior[:, :, trial] = scipy.interpolate.RectBivariateSpline(z, t, csd[:, :, trial])(xshort, t)
        lfp[:, :, trial] = fwd_model_1d(csd[:, :, trial], z, x, true["R"])
    lfp = lfp + np.random.normal(0, np.sqrt(true["sig2n"]), size=lfp.shape)
    lfp = normalize(lfp)
    test_lfp, test_csd = lfp[:, :, ntrials:], normalize(csd_interior[:, :, ntrials:])
    tcsd = normalize(predictcsd_trad_1d(test_lfp)[1:-1, :, :])

    model = GPCSD1D(test_lfp, x, t)
    set_true(model)
    assert "GPCSD1D object" in str(model)
    model.predict(xshort, t)
    gp = normalize(model.csd_pred)
    r2_gp, r2_t = _r2(gp, test_csd), _r2(tcsd, test_csd)
    assert model.csd_pred.shape == (nx - 2, nt, ntrials) and len(model.csd_pred_list) == 2
    assert r2_gp > 0.9 and r2_gp > r2_t                      # GPCSD recovers the CSD and beats tCSD

    # fit branch of the script (short): train on the first half, update_lfp to the second, predict
    np.random.seed(3)
    fitted = GPCSD1D(lfp[:, :, :ntrials], x, t, a=np.min(z), b=np.max(z))
    fitted.fit(n_restarts=4, options={'maxiter': 100, 'disp': False, 'gtol': 1e-5, 'ftol': 1e7 * np.finfo(float).eps})
    params = fitted.extract_model_params()
    assert np.isfinite(params['R']) and fitted.R['min'] <= params['R'] <= fitted.R['max']
    fitted.update_lfp(test_lfp, t)
    fitted.predict(xshort, t)
    assert _r2(normalize(fitted.csd_pred), test_csd) > 0.8


def test_nonfinite_and_error_semantics(cuda_lib):
    """np.seterr(all='ignore') semantics: NaN in the data flows through as a value; wrong shapes raise."""
    from gpcsd_b200.gpcsd1d import GPCSD1D
    np.random.seed(0)
    x = np.linspace(0, 2300, 24)[:, None]
    t = np.linspace(0, 40, 40)[:, None]
    lfp = np.random.randn(24, 40, 3)
    m = GPCSD1D(lfp, x, t)
    assert np.isfinite(m.loglik())
    bad = lfp.copy()
    bad[3, 5, 1] = np.nan
    m.update_lfp(bad, t)
    assert np.isnan(m.loglik())
    m.update_lfp(lfp, t)
    with pytest.raises(ValueError):
        m.predict(x, t[:-3])                                  # len(t*) != len(t): same failure mode as the reference
    with pytest.raises(ValueError):
        m.predict(x, t, type="nonsense")
    m.update_lfp(np.random.randn(23, 40, 3), t)
    with pytest.raises(ValueError):
        m.loglik()


def test_sim_from_gp_2d_flow(cuda_lib):
    """simulation_studies/sim_from_gp_2D.py:19-100 on a smaller dense grid: sample a 2-D CSD from the prior on a dense
    grid, forward-model it to a sparse electrode grid, switch the model to the electrode geometry with
    update_lfp(.., x=...), predict CSD on the dense grid and LFP at the electrodes with the true parameters."""
    from gpcsd_b200.covariances import GPCSDTemporalCovMatern, GPCSDTemporalCovSE
    from gpcsd_b200.forward_models import fwd_model_2d
    from gpcsd_b200.gpcsd2d import GPCSD2D
    from gpcsd_b200.utility_functions import expand_grid
    np.random.seed(0)
    a1, b1, a2, b2, nt = 0, 60, 0, 600, 10
    t = np.linspace(0, 100, nt)[:, None]
    nx1, nx2, nz1, nz2 = 4, 20, 8, 60
    x1, x2 = np.linspace(a1, b1, nx1)[:, None], np.linspace(a2, b2, nx2)[:, None]
    z1, z2 = np.linspace(a1, b1, nz1)[:, None], np.linspace(a2, b2, nz2)[:, None]
    x_grid, z_grid = expand_grid(x1, x2), expand_grid(z1, z2)
    gen = GPCSD2D(np.zeros((z_grid.shape[0], nt, 1)), x=z_grid, t=t, a1=a1, b1=b1, a2=a2, b2=b2,
                  temporal_cov_list=[GPCSDTemporalCovSE(t), GPCSDTemporalCovMatern(t)], ngl1=12, ngl2=40, eps=10.0)
    gen.R['value'] = 30.0
    gen.sig2n['value'] = 0.05
    gen.spatial_cov.params['ell1']['value'], gen.spatial_cov.params['ell2']['value'] = 40.0, 100.0
    gen.temporal_cov_list[0].params['ell']['value'], gen.temporal_cov_list[0].params['sigma2']['value'] = 5, 20
    gen.temporal_cov_list[1].params['ell']['value'], gen.temporal_cov_list[1].params['sigma2']['value'] = 1, 10
    csd_dense, nan_lfp = gen.sample_prior(1, type="csd")
    assert csd_dense.shape == (nz1 * nz2, nt, 1) and np.all(np.isnan(nan_lfp))
    clean = np.atleast_3d(fwd_model_2d(csd_dense.reshape((nz1, nz2, nt, -1)), z1, z2, x_grid, 30.0, gen.eps))
    lfp_sparse = clean + np.random.normal(0, np.sqrt(0.05), clean.shape)
    gen.update_lfp(lfp_sparse, t, x_grid)                       # geometry switch: dense CSD grid -> electrode grid
    assert gen.spatial_cov.x is x_grid
    gen.predict(z_grid, t, type="csd")
    gen.predict(x_grid, t, type="lfp")
    assert gen.csd_pred.shape == (nz1 * nz2, nt, 1) and gen.lfp_pred.shape == (nx1 * nx2, nt, 1)
    r2_lfp = _r2(gen.lfp_pred, clean)
    corr = np.corrcoef(gen.csd_pred.ravel(), csd_dense.ravel())[0, 1]
    print("\n[2-D flow] R2(lfp_pred, noiseless lfp) = %.3f, corr(csd_pred, true csd) = %.3f" % (r2_lfp, corr))
    assert r2_lfp > 0.9          # the posterior mean de-noises the LFP
    assert corr > 0.5            # and recovers the CSD pattern from 4 electrode columns
