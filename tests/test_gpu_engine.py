"""End-to-end parity of the CUDA engine against the oracle and the reference-generated golden vectors.

Tolerances (BASELINE.json north_star): relative 1e-9 on log-likelihood and gradient, 1e-8 on predicted
CSD/LFP.  Inputs are model-matched synthetic draws."""
import os

import numpy as np
import pytest

from helpers import engine_from_oracle, relerr, scaled_rel, ulp_sensitivity

pytestmark = pytest.mark.gpu

TOL_LL = 1e-9
TOL_GRAD = 1e-9
TOL_PRED = 1e-8


def solver_spread(fn):
    """Spread of the REFERENCE FORMULA's own result when only the LAPACK eigh driver changes
    (syevd vs syevr vs syev): the floor below which "parity" is not defined."""
    from oracle import gpcsd_oracle as O
    vals = [np.atleast_1d(fn(O.eigh_driver(d))) for d in ("evd", "evr", "ev")]
    ref = np.abs(vals[0])
    return max(float(np.max(np.abs(v - vals[0]) / ref)) for v in vals[1:])


def grad_tol(om, lfp):
    from oracle import gpcsd_oracle as O
    return max(TOL_GRAD, 10.0 * solver_spread(lambda e: O.loglik_and_grad(om, lfp, eigh=e)[1]))


def vector_noise_grad_tol(om, lfp):
    """Gate for the per-electrode-noise gradient: north_star's 1e-9, or 4 x the MEASURED sensitivity of the reference
    formula's own result to 1-ulp perturbations of the matrices it hands to eigh (helpers.ulp_sensitivity) where that is
    larger -- the function depends on eigenvector identity (utility_functions.py:54-57)."""
    sens = ulp_sensitivity(om, lfp)
    return max(TOL_GRAD, 4.0 * sens), sens


def _model_from_golden_1d(g):
    from oracle import gpcsd_oracle as O
    sp = O.Spatial1D(g["x"], float(g["a"]), float(g["b"]), int(g["ngl"]))
    temporal = [(int(k), float(e), float(s)) for k, e, s in zip(g["t_kind"], g["t_ell"], g["t_sigma2"])]
    sig = g["sig2n"]
    sig = float(sig) if sig.ndim == 0 else np.array(sig)
    return O.Model(1, sp, g["t"], float(g["R"]), (float(g["ell"]),), temporal, sig)


def _model_from_golden_2d(g):
    from oracle import gpcsd_oracle as O
    sp = O.Spatial2D(g["x"], float(g["a1"]), float(g["b1"]), float(g["a2"]), float(g["b2"]), int(g["ngl1"]), int(g["ngl2"]))
    temporal = [(int(k), float(e), float(s)) for k, e, s in zip(g["t_kind"], g["t_ell"], g["t_sigma2"])]
    return O.Model(2, sp, g["t"], float(g["R"]), (float(g["ell1"]), float(g["ell2"])), temporal, float(g["sig2n"]), float(g["eps"]))


@pytest.mark.parametrize("name", ["gpcsd1d_cfg1", "gpcsd1d_lownoise", "gpcsd1d_vecnoise"])
def test_loglik_matches_reference_golden_1d(cuda_lib, golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    om = _model_from_golden_1d(g)
    eng, hp = engine_from_oracle(om, g["lfp"])
    ll = eng.loglik(hp)
    assert abs(ll - float(g["loglik"])) / abs(float(g["loglik"])) < TOL_LL


def test_loglik_matches_reference_golden_2d(cuda_lib, golden_dir):
    g = np.load(os.path.join(golden_dir, "gpcsd2d_small.npz"))
    om = _model_from_golden_2d(g)
    eng, hp = engine_from_oracle(om, g["lfp"])
    ll = eng.loglik(hp)
    assert abs(ll - float(g["loglik"])) / abs(float(g["loglik"])) < TOL_LL


@pytest.mark.parametrize("name", ["gpcsd1d_cfg1", "gpcsd1d_lownoise"])
def test_grad_matches_oracle_and_reference_fd_1d(cuda_lib, golden_dir, name):
    from oracle import gpcsd_oracle as O
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    om = _model_from_golden_1d(g)
    eng, hp = engine_from_oracle(om, g["lfp"])
    ll, grad = eng.loglik_grad(hp)
    ll_o, grad_o = O.loglik_and_grad(om, g["lfp"])
    assert abs(ll - ll_o) / abs(ll_o) < TOL_LL
    # 1e-9, or 10x the reference formula's own LAPACK-driver spread where that is larger (the sig2n
    # component at sig2n = 1e-4 cancels two 4e7-sized terms down to 3e3)
    assert np.max(np.abs(grad - grad_o) / np.abs(grad_o)) < grad_tol(om, g["lfp"])
    # FD of the reference's own loglik: limited by FD truncation/roundoff, not by us
    fd = g["grad_fd_natural"]
    assert np.max(np.abs(grad - fd) / np.maximum(np.abs(fd), 1e-3 * np.max(np.abs(fd)))) < 1e-4


def test_grad_at_perturbed_point_vs_torch_autograd(cuda_lib, golden_dir):
    """Gradient evaluation point theta_true + 0.1 N(0,1); checker = torch autograd through eigh (the
    published reverse-mode algorithm the reference's autograd.grad implements)."""
    from oracle import gpcsd_oracle as O
    from oracle.oracle_torch import loglik_and_grad_torch
    g = np.load(os.path.join(golden_dir, "gpcsd1d_cfg1.npz"))
    om = _model_from_golden_1d(g)
    v = g["pert_natural"]
    om2 = O.Model(1, om.spatial, om.t, float(v[0]), (float(v[1]),), [(0, float(v[2]), float(v[3])), (1, float(v[4]), float(v[5]))], float(v[6]))
    eng, hp = engine_from_oracle(om2, g["lfp"])
    ll, grad = eng.loglik_grad(hp)
    assert abs(ll - float(g["pert_loglik"])) / abs(float(g["pert_loglik"])) < TOL_LL
    ll_t, grad_t = loglik_and_grad_torch(om2, g["lfp"])
    assert np.max(np.abs(grad - grad_t) / np.abs(grad_t)) < TOL_GRAD


def test_grad_2d(cuda_lib, golden_dir):
    from oracle import gpcsd_oracle as O
    g = np.load(os.path.join(golden_dir, "gpcsd2d_small.npz"))
    om = _model_from_golden_2d(g)
    eng, hp = engine_from_oracle(om, g["lfp"])
    ll, grad = eng.loglik_grad(hp)
    ll_o, grad_o = O.loglik_and_grad(om, g["lfp"])
    assert abs(ll - ll_o) / abs(ll_o) < TOL_LL
    assert np.max(np.abs(grad - grad_o) / np.abs(grad_o)) < TOL_GRAD


def test_grad_vector_noise_end_to_end(cuda_lib, golden_dir):
    """Per-electrode noise enters by spatial EIGEN index (util:54-57) so loglik depends on eigenvector
    identity; end-to-end parity (own eigensolver vs numpy eigh inside the oracle's closed form)."""
    from oracle import gpcsd_oracle as O
    g = np.load(os.path.join(golden_dir, "gpcsd1d_vecnoise.npz"))
    om = _model_from_golden_1d(g)
    eng, hp = engine_from_oracle(om, g["lfp"])
    ll, grad = eng.loglik_grad(hp)
    ll_o, grad_o = O.loglik_and_grad(om, g["lfp"])
    assert len(grad) == 6 + 24
    assert abs(ll - ll_o) / abs(ll_o) < TOL_LL
    # eigenvector-identity conditioning: the gate follows the measured 1-ulp sensitivity of the
    # reference formula on these inputs, not a constant; tests/test_gpu_factor_parity.py holds the identical-factor check
    tol, sens = vector_noise_grad_tol(om, g["lfp"])
    rel = scaled_rel(grad, grad_o)
    print("\n[vector noise, golden] grad rel max %.2e; reference formula's 1-ulp sensitivity %.2e -> gate %.1e" % (rel.max(), sens, tol))
    assert rel.max() < tol


@pytest.mark.parametrize("name,is2d", [("gpcsd1d_cfg1", False), ("gpcsd1d_lownoise", False), ("gpcsd2d_small", True)])
def test_predict_matches_oracle_and_golden(cuda_lib, golden_dir, name, is2d):
    from oracle import gpcsd_oracle as O
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    om = _model_from_golden_2d(g) if is2d else _model_from_golden_1d(g)
    eng, hp = engine_from_oracle(om, g["lfp"])
    out = eng.predict(hp, g["z"], g["t"], "both")
    ref = O.predict_kron(om, g["lfp"], g["z"], g["t"], "both")
    npred = g["csd_pred"].shape[2]
    for key in ("csd_pred", "lfp_pred"):
        assert out[key].shape == ref[key].shape
        assert relerr(out[key], ref[key]) < TOL_PRED
        for k in range(2):
            assert relerr(out[key + "_list"][k], ref[key + "_list"][k]) < TOL_PRED
        # reference's dense (nx nt)^2 inverse: its own conditioning limits agreement
        assert relerr(out[key][:, :, :npred], g[key]) < 1e-6
        assert relerr(out[key + "_list"][1][:, :, :npred], g[key + "_1"]) < 1e-6


def test_predict_requires_matching_time_grid(cuda_lib, golden_dir):
    g = np.load(os.path.join(golden_dir, "gpcsd1d_lownoise.npz"))
    om = _model_from_golden_1d(g)
    eng, hp = engine_from_oracle(om, g["lfp"])
    with pytest.raises(ValueError):
        eng.predict(hp, g["z"], g["t"][:-1], "csd")


def test_single_trial_and_ragged_trial_counts(cuda_lib):
    from oracle import gpcsd_oracle as O, synth
    x, t = synth.geometry_1d(24, 33)
    om = synth.model_1d(x, t)
    for N in (1, 7, 129):
        lfp = synth.matched_lfp(om, N, 50 + N)
        eng, hp = engine_from_oracle(om, lfp)
        ll, grad = eng.loglik_grad(hp)
        ll_o, grad_o = O.loglik_and_grad(om, lfp)
        assert abs(ll - ll_o) / abs(ll_o) < TOL_LL
        assert np.max(np.abs(grad - grad_o) / np.abs(grad_o)) < TOL_GRAD


@pytest.mark.parametrize("nt,uniform,fold", [(41, True, False), (64, True, False), (40, False, False), (41, True, True),
                                             (64, True, True), (131, True, True), (130, True, True)])
def test_temporal_eigh_paths_agree_with_oracle(cuda_lib, nt, uniform, fold):
    """Uniform grid (even / odd nt) -> centrosymmetric split, and (nt >= FOLD_MIN_NT, or forced here) the folded time basis
    for the projection, the temporal SYRK and predict's back-projection; jittered grid -> one solve of order nt."""
    from gpcsd_b200.engine import KronEngine
    old_min = KronEngine.FOLD_MIN_NT
    KronEngine.FOLD_MIN_NT = 32 if fold else 10 ** 9
    try:
        _temporal_paths(nt, uniform)
    finally:
        KronEngine.FOLD_MIN_NT = old_min


def _temporal_paths(nt, uniform):
    from oracle import gpcsd_oracle as O, synth
    x, t = synth.geometry_1d(24, nt)
    if not uniform:
        t = t + 0.05 * np.sin(np.arange(nt))[:, None]
    om = synth.model_1d(x, t)
    lfp = synth.matched_lfp(om, 9, 5 + nt)
    eng, hp = engine_from_oracle(om, lfp)
    assert eng.t_uniform == uniform
    ll, grad = eng.loglik_grad(hp)
    ll_o, grad_o = O.loglik_and_grad(om, lfp)
    assert abs(ll - ll_o) / abs(ll_o) < TOL_LL
    assert np.max(np.abs(grad - grad_o) / np.abs(grad_o)) < grad_tol(om, lfp)
    out = eng.predict(hp, x, t, "csd")
    ref = O.predict_kron(om, lfp, x, t, "csd")
    assert relerr(out["csd_pred"], ref["csd_pred"]) < TOL_PRED


def test_spatial_reflection_symmetry_split(cuda_lib):
    """Neuropixels-like checkerboard with symmetric integration bounds: the point reflection about the box centre maps
    the sites onto themselves, the engine splits the spatial eigenproblem, and results still match the oracle.  A shifted
    box (no symmetry) and per-electrode noise (eigenvalue order needed) must take the unsplit path."""
    from oracle import gpcsd_oracle as O, synth
    X, t = synth.geometry_neuropixels(96, 40, 0.4)
    ymax = X[:, 1].max()
    om = synth.model_2d(X, t, ngl1=10, ngl2=40, a1=-16.0, b1=64.0, a2=-100.0, b2=ymax + 100.0, eps=1.0, sig2n=0.5)
    lfp = synth.matched_lfp(om, 7, 12)
    eng, hp = engine_from_oracle(om, lfp)
    assert eng.s_pairs is not None and eng.s_pairs[0].numel() == 48
    ll, grad = eng.loglik_grad(hp)
    ll_o, grad_o = O.loglik_and_grad(om, lfp)
    assert abs(ll - ll_o) / abs(ll_o) < TOL_LL
    assert np.max(np.abs(grad - grad_o) / np.abs(grad_o)) < grad_tol(om, lfp)
    out = eng.predict(hp, X[::7], t, "both")
    ref = O.predict_kron(om, lfp, X[::7], t, "both")
    assert relerr(out["csd_pred"], ref["csd_pred"]) < TOL_PRED and relerr(out["lfp_pred"], ref["lfp_pred"]) < TOL_PRED
    # a new upload must refresh the channel-folded copy of the LFP
    eng.set_lfp(0.5 * lfp[:, :, ::-1].copy())
    ll2_o = O.loglik(om, 0.5 * lfp[:, :, ::-1])
    assert abs(eng.loglik(hp) - ll2_o) / abs(ll2_o) < TOL_LL
    # shifted box: not symmetric
    om_shift = synth.model_2d(X, t, ngl1=10, ngl2=40, a1=-16.0, b1=64.0, a2=-100.0, b2=ymax + 140.0, eps=1.0, sig2n=0.5)
    eng2, hp2 = engine_from_oracle(om_shift, lfp)
    assert eng2.s_pairs is None
    assert abs(eng2.loglik(hp2) - O.loglik(om_shift, lfp)) / abs(O.loglik(om_shift, lfp)) < TOL_LL
    # 1-D probe with symmetric bounds and per-electrode noise: symmetry detected but NOT used (ascending order matters)
    x1, t1 = synth.geometry_1d(24, 40)
    rng = np.random.default_rng(0)
    om1 = synth.model_1d(x1, t1, sig2n=1e-2 * np.exp(0.3 * rng.standard_normal(24)))
    lfp1 = synth.matched_lfp(om1, 5, 3)
    eng3, hp3 = engine_from_oracle(om1, lfp1)
    assert eng3.s_pairs is not None
    assert abs(eng3.loglik(hp3) - O.loglik(om1, lfp1)) / abs(O.loglik(om1, lfp1)) < 1e-8


def test_predict_against_multiprecision_arbiter(cuda_lib):
    """Engine predict vs a 40-digit mpmath solve of the same system (oracle/arbiter_mp.py) at cond(K) ~ 6e8: the engine's
    Kronecker form agrees with the arbiter far below the 1e-8 gate; tests/test_predict_arbiter.py shows on the CPU that the
    reference's dense formulation is the less accurate side when the two float64 results differ."""
    from oracle import synth
    from oracle.arbiter_mp import predict_arbiter
    x, t = synth.geometry_1d(10, 14)
    om = synth.model_1d(x, t, sig2n=1e-7)
    lfp = synth.matched_lfp(om, 2, 5)
    z = np.linspace(100.0, 2200.0, 5)[:, None]
    eng, hp = engine_from_oracle(om, lfp)
    out = eng.predict(hp, z, t, "csd")
    arb = predict_arbiter(om, lfp, z, "csd")
    for c in range(2):
        err = relerr(out["csd_pred_list"][c], arb[c])
        print("\n[predict vs 40-digit arbiter] component %d: %.2e" % (c, err))
        assert err < 1e-10
