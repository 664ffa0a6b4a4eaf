"""CPU: the oracle (numpy restatement) against the golden vectors generated from the unmodified reference
(oracle/make_golden.py).  Forward quantities must match to round-off; gradients against 4th-order FD of the
reference's own loglik."""
import os

import numpy as np
import pytest

from helpers import relerr
from test_gpu_engine import _model_from_golden_1d, _model_from_golden_2d


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


@pytest.mark.parametrize("name", ["gpcsd1d_cfg1", "gpcsd1d_lownoise", "gpcsd1d_vecnoise"])
def test_oracle_1d_forward(golden_dir, name):
    from oracle import gpcsd_oracle as O
    g = _load(golden_dir, name)
    om = _model_from_golden_1d(g)
    assert relerr(om.spatial.gl_x, g["gl_x"]) < 1e-15 and relerr(om.spatial.gl_w, g["gl_w"]) < 1e-15
    assert relerr(om.Ks(), g["Ks"]) < 1e-14
    assert relerr(om.Kphig(g["z"]), g["Kphig"]) < 1e-14
    assert relerr(om.Ks(xp=g["z"]), g["Kphi_z"]) < 1e-14
    assert relerr(O.compute_Ks_1d(g["x"], float(g["ell"])), g["Ks_csd"]) < 1e-15
    (k0, e0, s0), (k1, e1, s1) = om.temporal
    assert relerr(O.compute_Kt(k0, e0, s0, g["t"]), g["Kt_se"]) < 1e-15
    assert relerr(O.compute_Kt(k1, e1, s1, g["t"]), g["Kt_matern"]) < 1e-15
    ll = O.loglik(om, g["lfp"])
    assert abs(ll - float(g["loglik"])) <= 1e-12 * abs(float(g["loglik"]))
    # the kernel-level restatement (einsum over all trials) agrees with the literal trial loop
    Qs, Qt, D, ls, lt = O.comp_eig_D(om.Ks(jitter=True), om.Kt(), om.sig2n)
    assert abs(O.loglik_from_factors(g["lfp"], Qs, ls, Qt, lt, om.sig2n) - ll) <= 1e-11 * abs(ll)


@pytest.mark.parametrize("name,is2d", [("gpcsd1d_cfg1", False), ("gpcsd1d_lownoise", False), ("gpcsd1d_vecnoise", False), ("gpcsd2d_small", True)])
def test_oracle_predict(golden_dir, name, is2d):
    from oracle import gpcsd_oracle as O
    g = _load(golden_dir, name)
    om = _model_from_golden_2d(g) if is2d else _model_from_golden_1d(g)
    npred = g["csd_pred"].shape[2]
    lfp = g["lfp"][:, :, :npred]
    dense = O.predict_dense(om, lfp, g["z"], g["t"], "both")
    kron = O.predict_kron(om, lfp, g["z"], g["t"], "both")
    for key in ("csd_pred", "lfp_pred"):
        assert relerr(dense[key], g[key]) < 1e-9          # literal restatement (dense inverse) == reference
        assert relerr(dense[key + "_list"][0], g[key + "_0"]) < 1e-9
        assert relerr(dense[key + "_list"][1], g[key + "_1"]) < 1e-9
        # Kronecker form == dense form up to the dense inverse's own conditioning
        assert relerr(kron[key], g[key]) < 1e-6


def test_oracle_2d_forward(golden_dir):
    from oracle import gpcsd_oracle as O
    g = _load(golden_dir, "gpcsd2d_small")
    om = _model_from_golden_2d(g)
    sp = om.spatial
    assert relerr(np.stack([sp.grid1, sp.grid2], 1), g["gl_x_grid"]) < 1e-15
    assert relerr(sp.w_prod, g["gl_w_prod"].reshape(-1)) < 1e-15
    assert relerr(om.Ks(), g["Ks"]) < 1e-14
    assert relerr(om.Kphig(g["z"]), g["Kphig"]) < 1e-14
    assert relerr(om.Ks(xp=g["z"]), g["Kphi_z"]) < 1e-14
    assert relerr(O.compute_Ks_2d(g["x"], float(g["ell1"]), float(g["ell2"])), g["Ks_csd"]) < 1e-15
    ll = O.loglik(om, g["lfp"])
    assert abs(ll - float(g["loglik"])) <= 1e-12 * abs(float(g["loglik"]))


@pytest.mark.parametrize("name,is2d", [("gpcsd1d_cfg1", False), ("gpcsd1d_lownoise", False), ("gpcsd2d_small", True)])
def test_oracle_gradient_vs_reference_fd_and_autograd(golden_dir, name, is2d):
    from oracle import gpcsd_oracle as O
    from oracle.oracle_torch import loglik_and_grad_torch
    g = _load(golden_dir, name)
    om = _model_from_golden_2d(g) if is2d else _model_from_golden_1d(g)
    ll, grad = O.loglik_and_grad(om, g["lfp"])
    fd = g["grad_fd_natural"]
    # FD of the reference: truncation + cancellation limited (1e-4 relative step, 4th order)
    assert np.max(np.abs(grad - fd) / np.maximum(np.abs(fd), 1e-3 * np.max(np.abs(fd)))) < 1e-4
    ll_t, grad_t = loglik_and_grad_torch(om, g["lfp"])
    assert abs(ll - ll_t) <= 1e-11 * abs(ll)
    assert np.max(np.abs(grad - grad_t) / np.abs(grad_t)) < 1e-7


def test_oracle_obj_fun_chain_rule(golden_dir):
    """nll gradient in log space (gpcsd1d.py:160-174 transforms + priors) against central differences."""
    from oracle import gpcsd_oracle as O, synth
    g = _load(golden_dir, "gpcsd1d_lownoise")
    om = _model_from_golden_1d(g)
    pri = synth.default_priors(om)
    tp = O.pack_tparams(om)
    f0, gr = O.obj_and_grad(om, g["lfp"], tp, pri)
    assert abs(f0 - O.obj_fun(om, g["lfp"], tp, pri)) < 1e-9 * abs(f0)
    for k in range(len(tp)):
        e = np.zeros_like(tp)
        e[k] = 1e-5
        fd = (O.obj_fun(om, g["lfp"], tp + e, pri) - O.obj_fun(om, g["lfp"], tp - e, pri)) / 2e-5
        assert abs(fd - gr[k]) < 2e-5 * max(abs(gr[k]), 1.0), (k, fd, gr[k])


def test_helpers_golden(golden_dir):
    """forward models, mykron, comp_eig_D, priors, grids, tCSD restatements vs reference outputs."""
    from oracle import gpcsd_oracle as O
    from gpcsd_b200 import forward_models as fm, predict_csd as pc, priors as pr, utility_functions as uf
    g = _load(golden_dir, "helpers")
    assert relerr(O.b_fwd_1d(g["r"], 80.0), g["b1d"]) < 1e-15
    assert relerr(fm.b_fwd_1d(g["r"], 80.0), g["b1d"]) < 1e-15
    assert relerr(fm.b_fwd_2d(g["r"], 0.5 * g["r"], 80.0, 20.0), g["b2d"]) < 1e-15
    assert relerr(O.mykron(g["A"], g["B"]), g["kron"]) == 0.0
    assert relerr(uf.mykron(g["A"], g["B"]), g["kron"]) == 0.0
    _, _, D, _, _ = O.comp_eig_D(g["Ks"], g["Kt"], 0.3)
    assert relerr(D, g["Dvec"]) < 1e-13
    _, _, Dv, _, _ = O.comp_eig_D(g["Ks"], g["Kt"], g["sv"])
    assert relerr(Dv, g["Dvec_vec"]) < 1e-13
    assert relerr(O.fwd_model_1d(g["csd"], g["xd"], g["zz"], 120.0, varsigma=0.4), g["fwd1d"]) < 1e-13
    # (gpcsd_b200.forward_models.fwd_model_1d/2d run on the device: tests/test_gpu_next_rows.py holds their golden check)
    assert np.array_equal(pc.predictcsd_trad_1d(g["lf"]), g["tcsd1"])
    assert np.array_equal(pc.predictcsd_trad_2d(g["lf4"]), g["tcsd2"], equal_nan=True)
    ig = pr.GPCSDInvGammaPrior()
    ig.set_params(3.0, 40.0)
    assert ig.alpha == float(g["ig_alpha"]) and ig.beta == float(g["ig_beta"])
    assert relerr([ig.lpdf(v) for v in g["xs"]], g["ig_lpdf"]) < 1e-15
    assert relerr([pr.GPCSDHalfNormalPrior(0.7).lpdf(v) for v in g["xs"]], g["hn_lpdf"]) < 1e-15
    assert ig.lpdf(-1.0) == -np.inf and pr.GPCSDHalfNormalPrior(0.7).lpdf(0.0) == -np.inf
    assert np.array_equal(uf.expand_grid(g["x1"], g["x2"]), g["grid"])
    assert np.array_equal(uf.sort_grid(g["grid_perm"]), g["grid_sorted"])
    assert np.array_equal(uf.normalize(g["norm_in"]), g["norm_out"])
    a, b = uf.reduce_grid(g["grid"])
    assert np.array_equal(a, g["x1"].reshape(-1)) and np.array_equal(b, g["x2"].reshape(-1))
    # prior derivative helpers (closed form of priors.py:27, :50)
    for p in (ig, pr.GPCSDHalfNormalPrior(0.7)):
        for v in g["xs"]:
            fd = (p.lpdf(v + 1e-6) - p.lpdf(v - 1e-6)) / 2e-6
            assert abs(fd - p.dlpdf(v)) < 1e-6 * max(1.0, abs(fd))


def test_reference_shim_agrees_when_available():
    """If the reference tree is mounted (build container), the oracle must still reproduce it live."""
    from oracle import ref_shim
    if not ref_shim.reference_available():
        pytest.skip("reference tree not present (GPU box)")
    import subprocess, sys, os
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r)
from oracle.ref_shim import import_reference
from oracle import gpcsd_oracle as O, synth
g = import_reference()
x, t = synth.geometry_1d(24, 30)
om = synth.model_1d(x, t, sig2n=3e-3)
lfp = synth.matched_lfp(om, 4, 77)
np.random.seed(0)
m = g.gpcsd1d.GPCSD1D(lfp, x, t)
m.R['value'] = om.R; m.spatial_cov.params['ell']['value'] = om.ells[0]
for tc, (_, e, s) in zip(m.temporal_cov_list, om.temporal):
    tc.params['ell']['value'] = e; tc.params['sigma2']['value'] = s
m.sig2n['value'] = om.sig2n
assert abs(float(m.loglik()) - O.loglik(om, lfp)) < 1e-12 * abs(O.loglik(om, lfp))
m.predict(x, t, type='csd')
assert np.max(np.abs(O.predict_dense(om, lfp, x, t, 'csd')['csd_pred'] - m.csd_pred)) < 1e-9 * np.max(np.abs(m.csd_pred))
print('ok')
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]
