"""-m gpu: the callers either side of the hot path on the device, each against the oracle / the
reference-generated golden vectors: forward-model operators, Cholesky + Philox normals + sample_prior, and the per-trial
evoked-shift objective with its batched fit."""
import os

import numpy as np
import pytest

from helpers import engine_from_oracle, relerr

pytestmark = pytest.mark.gpu


# ---------------------------------------------------------------------------------------------------- forward models
def test_fwd_models_match_reference_golden(cuda_lib, golden_dir):
    from gpcsd_b200 import forward_models as fm
    g = np.load(os.path.join(golden_dir, "helpers.npz"))
    assert relerr(fm.fwd_model_1d(g["csd"], g["xd"], g["zz"], 120.0, varsigma=0.4), g["fwd1d"]) < 1e-13
    assert relerr(fm.fwd_model_2d(g["arr2"], g["x1"], g["x2"], g["z2"], 60.0, 10.0), g["fwd2d"]) < 1e-13


def test_fwd_model_1d_nonuniform_grid_vs_oracle(cuda_lib):
    from gpcsd_b200 import forward_models as fm
    from oracle import gpcsd_oracle as O
    rng = np.random.default_rng(1)
    xd = np.sort(rng.uniform(0.0, 2400.0, 301))[:, None]          # non-uniform integration grid
    z = np.linspace(50.0, 2300.0, 37)[:, None]
    csd = rng.standard_normal((301, 45))
    assert relerr(fm.fwd_model_1d(csd, xd, z, 150.0, varsigma=0.3), O.fwd_model_1d(csd, xd, z, 150.0, varsigma=0.3)) < 1e-13


# ---------------------------------------------------------------------------------------------------- Cholesky / RNG
@pytest.mark.parametrize("n", [1, 7, 24, 33, 64, 200, 1000])
def test_cholesky_matches_numpy(cuda_lib, n):
    from gpcsd_b200 import devops
    t = np.linspace(0.0, 3.0, n)[:, None]
    K = np.exp(-0.5 * np.square((t - t.T) / 0.7)) + 0.7 * np.exp(-np.abs(t - t.T) / 0.2) + 1e-6 * np.eye(n)
    Lh = devops.cholesky(K)
    ref = np.linalg.cholesky(K)
    assert np.allclose(np.triu(Lh, 1), 0.0)
    assert relerr(Lh @ Lh.T, K) < 1e-13
    assert relerr(Lh, ref) < 1e-9                      # cond(K) ~ 1e6: factor entries differ at eps * cond


def test_cholesky_not_positive_definite_raises(cuda_lib):
    from gpcsd_b200 import devops
    K = np.eye(40)
    K[17, 17] = -1.0
    with pytest.raises(np.linalg.LinAlgError):
        devops.cholesky(K)


def test_philox_bits_and_normals_match_oracle(cuda_lib):
    import torch
    from gpcsd_b200 import _lib as L, devops
    from oracle import philox
    n = 5000
    out = torch.zeros(4 * n, dtype=torch.int32, device="cuda")
    L.call("gpcsd_philox_raw", n, 0x0123456789ABCDEF, 7, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    got = out.cpu().numpy().view(np.uint32).reshape(n, 4)
    assert np.array_equal(got, philox.raw(n, 0x0123456789ABCDEF, 7))            # bit-exact
    z = devops.randn((3, 1667), seed=99, stream_id=2)
    ref = philox.randn(3 * 1667, 99, 2).reshape(3, 1667)
    assert np.max(np.abs(z - ref)) < 1e-13                                       # libm vs CUDA log / sincospi rounding
    # padded layout: live columns only, padding untouched
    zp = devops.randn_device(5, 6, 99, 2, ld=8).cpu().numpy()
    assert np.max(np.abs(zp[:, :6].reshape(-1) - ref.reshape(-1)[:30])) < 1e-13 and np.all(zp[:, 6:] == 0.0)


def test_sample_prior_device_generator(cuda_lib):
    """device=True: Ls Z Lt^T with Z from the Philox stream -- checked against numpy with the oracle's copy of the stream --
    and its second moments against Ks (x) Kt."""
    from gpcsd_b200.gpcsd1d import GPCSD1D
    from oracle import gpcsd_oracle as O, philox
    np.random.seed(3)
    x = np.linspace(0, 2300, 24)[:, None]
    t = np.linspace(0, 30, 31)[:, None]
    m = GPCSD1D(np.zeros((24, 31, 1)), x, t)
    m.spatial_cov.params['ell']['value'] = 300.0
    N = 4000
    csd = m.sample_prior(N, device=True, seed=17)
    assert csd.is_cuda and tuple(csd.shape) == (24, 31, N)
    csd = csd.cpu().numpy()
    Kt = sum(O.compute_Kt(tc.KIND, tc.params['ell']['value'], tc.params['sigma2']['value'], t) for tc in m.temporal_cov_list)
    Ks = O.compute_Ks_1d(x, 300.0) + 1e-8 * np.eye(24)
    Lt, Ls = np.linalg.cholesky(Kt), np.linalg.cholesky(Ks)
    Z = philox.randn(24 * 31 * N, 17, 0).reshape(24, 31, N)
    ref = np.einsum("ia,ajr,bj->ibr", Ls, Z, Lt, optimize=True)
    assert relerr(csd, ref) < 1e-7                      # cond(Ks + 1e-8 I) ~ 1e8 (same gate as the host-RNG test)
    # second moments: E[csd_i,j csd_i',j] = Ks[i,i'] Kt[j,j]
    emp = np.einsum("ijr,kjr->ik", csd, csd) / (N * np.trace(Kt))
    assert relerr(emp, Ks) < 0.08


def test_sample_prior_2d_device_and_host(cuda_lib):
    from gpcsd_b200.gpcsd2d import GPCSD2D
    from oracle import synth
    X, t = synth.geometry_grid_2d(3, 6, 12)
    np.random.seed(1)
    m = GPCSD2D(np.zeros((18, 12, 1)), X, t, ngl1=6, ngl2=10)
    csd, lfp = m.sample_prior(6, type="both", seed=4)
    assert csd.shape == lfp.shape == (18, 12, 6) and np.all(np.isfinite(csd)) and np.all(np.isfinite(lfp))
    csd2, lfp2 = m.sample_prior(6, type="csd", seed=4)
    assert np.array_equal(csd, csd2) and np.all(np.isnan(lfp2))
    dcsd, dlfp = m.sample_prior(6, type="csd", seed=4, device=True)
    assert dcsd.is_cuda and dlfp is None and tuple(dcsd.shape) == (18, 12, 6)


# ---------------------------------------------------------------------------------------------------- evoked shifts
def _shift_case(nt, uniform, N, seed):
    from oracle import gpcsd_oracle as O, synth
    x, t = synth.geometry_1d(24, nt, ms_grid=True)
    if not uniform:
        t = t + 0.2 * np.sin(np.arange(nt))[:, None]
    rng = np.random.default_rng(seed)
    om = synth.model_1d(x, t, sig2n=1e-2 * np.exp(0.3 * rng.standard_normal(24)))      # per-electrode noise like the script
    tt = t.squeeze()
    mu = np.zeros((24, nt, 3))
    mu[:, :, 0] = 0.1 * rng.standard_normal((24, nt))
    mu[:, :, 1] = np.exp(-0.5 * ((tt - 0.3 * nt) / 4.0) ** 2)[None, :] * rng.standard_normal((24, 1))
    mu[:, :, 2] = np.exp(-0.5 * ((tt - 0.6 * nt) / 6.0) ** 2)[None, :] * rng.standard_normal((24, 1))
    tau_true = 2.0 * rng.standard_normal((N, 2))
    lfp = synth.matched_lfp(om, N, seed + 1)
    for r in range(N):
        for s in range(2):
            lfp[:, :, r] += np.stack([np.interp(tt + tau_true[r, s], tt, mu[i, :, s + 1]) for i in range(24)])
        lfp[:, :, r] += mu[:, :, 0]
    return om, lfp, mu, tau_true


@pytest.mark.parametrize("nt,uniform", [(60, True), (130, True), (131, True), (40, False)])
def test_shift_objective_matches_oracle(cuda_lib, nt, uniform):
    """nll and d nll / d tau for every trial vs the literal restatement of fit_mean_function.py:311-321 (identical factors on
    both sides: 1e-11) and vs the oracle with its own numpy eigh (1e-9); shifts include values that extrapolate."""
    from oracle import gpcsd_oracle as O
    N = 9
    om, lfp, mu, _ = _shift_case(nt, uniform, N, 3)
    rng = np.random.default_rng(8)
    tau = 3.0 * rng.standard_normal((N, 2))
    tau[0] = [0.0, 0.0]
    tau[1] = [nt * 0.7, -nt * 0.6]                                # far outside: both end intervals extrapolate
    eng, hp = engine_from_oracle(om, lfp)
    nll, grad = eng.shift_objective(hp, mu, tau)
    Qs, Qt, D, ls, lt = O.comp_eig_D(om.Ks(jitter=True), om.Kt(), om.sig2n)
    ref = np.array([O.shift_objective(lfp[:, :, r], mu, om.t, tau[r], Qs, Qt, D) for r in range(N)])
    gref = np.array([O.shift_objective_grad(lfp[:, :, r], mu, om.t, tau[r], Qs, Qt, D) for r in range(N)])
    # own eigensolver vs numpy eigh: per-electrode noise makes the value depend on eigenvector identity
    # (utility_functions.py:54-57), so the gate is 1e-9 or 4 x the MEASURED change of the oracle's own values under 1-ulp
    # perturbations of (Ks, Kt) before eigh, whichever is larger; the identical-factor comparison below is the kernel check
    sens = 0.0
    for seed in range(3):
        prng = np.random.default_rng(seed)
        pert = lambda K: K + np.finfo(float).eps * np.abs(K) * (lambda E: 0.5 * (E + E.T))(prng.uniform(-1, 1, K.shape))
        Qs2, Qt2, D2, _, _ = O.comp_eig_D(pert(om.Ks(jitter=True)), pert(om.Kt()), om.sig2n)
        v = np.array([O.shift_objective(lfp[:, :, r], mu, om.t, tau[r], Qs2, Qt2, D2) for r in range(N)])
        sens = max(sens, float(np.max(np.abs(v - ref) / np.abs(ref))))
    tol = max(1e-9, 4.0 * sens)
    err = float(np.max(np.abs(nll - ref) / np.abs(ref)))
    print("\n[shift objective nt=%d] rel err %.2e; oracle's 1-ulp sensitivity %.2e -> gate %.1e" % (nt, err, sens, tol))
    assert err < tol
    assert np.max(np.abs(grad - gref)) < max(1e-8, 40.0 * sens) * np.max(np.abs(gref))
    nll_f, grad_f = eng.shift_objective(hp, mu, tau, factors=(Qs, ls, Qt, lt))
    assert np.max(np.abs(nll_f - ref) / np.abs(ref)) < 1e-11
    assert np.max(np.abs(grad_f - gref)) < 1e-11 * np.max(np.abs(gref))


def test_per_trial_shift_fit_matches_scipy_per_trial(cuda_lib):
    """All trials optimised in lock step (one batched device evaluation per step) reach the optima scipy's L-BFGS-B finds
    trial by trial on the oracle objective (the reference's minfunc, fit_mean_function.py:323-325), and recover the shifts."""
    import scipy.optimize
    from gpcsd_b200.batched_opt import batched_lbfgsb
    from oracle import gpcsd_oracle as O
    N = 12
    om, lfp, mu, tau_true = _shift_case(80, True, N, 11)
    eng, hp = engine_from_oracle(om, lfp)

    def fun(T, idx):
        full = np.zeros((N, 2))
        full[idx] = T
        f, g = eng.shift_objective(hp, mu, full)
        return f[idx], g[idx]
    res = batched_lbfgsb(fun, np.zeros((N, 2)), maxiter=100, gtol=1e-6, ftol=1e-12)
    Qs, Qt, D, _, _ = O.comp_eig_D(om.Ks(jitter=True), om.Kt(), om.sig2n)
    for r in range(0, N, 4):
        ref = scipy.optimize.minimize(lambda tv: O.shift_objective(lfp[:, :, r], mu, om.t, tv, Qs, Qt, D), np.zeros(2),
                                      method="l-bfgs-b")
        assert abs(res["fun"][r] - ref.fun) < 1e-5 * abs(ref.fun)
        assert np.max(np.abs(res["x"][r] - ref.x)) < 5e-3
    assert np.max(np.abs(res["x"] - tau_true)) < 0.5          # the shifts that generated the data are recovered
