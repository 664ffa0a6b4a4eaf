"""CPU: host-side logic of the drop-in API (construction, dictionaries, transforms, bounds, priors, RNG
coupling with the reference) -- everything that needs no kernel."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT


def _model_1d(vec=False, seed=3):
    from gpcsd_b200.gpcsd1d import GPCSD1D
    from gpcsd_b200.priors import GPCSDHalfNormalPrior
    np.random.seed(seed)
    x = np.linspace(0, 2300, 24)[:, None]
    t = np.linspace(0, 50, 40)[:, None]
    pri = [GPCSDHalfNormalPrior(0.1) for _ in range(24)] if vec else None
    return GPCSD1D(np.random.randn(24, 40, 3), x, t, sig2n_prior=pri)


def _model_2d(seed=4):
    from gpcsd_b200.gpcsd2d import GPCSD2D
    from gpcsd_b200.utility_functions import expand_grid
    np.random.seed(seed)
    X = expand_grid(np.linspace(0, 48, 4), np.linspace(0, 220, 12))
    t = np.arange(30.0)[:, None]
    return GPCSD2D(np.random.randn(48, 30, 2), X, t, ngl1=6, ngl2=11)


def test_extract_restore_roundtrip_and_str():
    m = _model_1d()
    p = m.extract_model_params()
    assert set(p) == {"R", "sig2n", "spatial_ell", "temporal_ell_list", "temporal_sigma2_list"}
    m2 = _model_1d(seed=9)
    m2.restore_model_params(p)
    assert m2.extract_model_params() == p
    assert str(m).startswith("GPCSD1D object\nLFP shape: (24, 40, 3)")
    m2d = _model_2d()
    p2 = m2d.extract_model_params()
    assert set(p2) == {"R", "eps", "sig2n", "spatial_ell1", "spatial_ell2", "temporal_ell_list", "temporal_sigma2_list"}
    assert str(m2d).startswith("GPCSD1D object")          # sic, gpcsd2d.py:82
    bad = dict(p)
    bad["temporal_ell_list"] = [1.0]
    m2.restore_model_params(bad)                            # prints and returns (gpcsd1d.py:97-99)
    assert m2.extract_model_params()["temporal_ell_list"] == p["temporal_ell_list"]


@pytest.mark.parametrize("vec", [False, True])
def test_tparams_transform_bounds_and_prior_chain(vec):
    m = _model_1d(vec)
    b = m._bounds()
    nparam = 6 + (24 if vec else 1)
    assert len(b) == nparam
    assert b[0] == (np.log(m.R["min"] / 100), np.log(m.R["max"] / 100))
    assert b[5][0] == -np.inf and b[5][1] == np.inf        # Matern sigma2 in [0, inf) (covariances.py:289)
    tp = np.linspace(-0.3, 0.4, nparam)
    m._set_tparams(tp, fix_R=False)
    assert np.isclose(m.R["value"], np.exp(tp[0]) * 100) and np.isclose(m.spatial_cov.params["ell"]["value"], np.exp(tp[1]) * 100)
    assert np.isclose(m.temporal_cov_list[1].params["sigma2"]["value"], np.exp(tp[5]))
    if vec:
        assert np.allclose(m.sig2n["value"], np.exp(tp[6:]))
    else:
        assert np.isclose(m.sig2n["value"], np.exp(tp[6]))
    R0 = m.R["value"]
    m._set_tparams(tp + 1.0, fix_R=True)
    assert m.R["value"] == R0
    lp, dlp, vals = m._prior_terms()
    assert len(dlp) == nparam and np.isfinite(lp)
    hp = m._hyperparams()
    assert hp.vector_noise == vec and hp.n_params() == nparam
    t0 = m._sample_tparams0(fix_R=True)
    assert np.isclose(t0[0], np.log(m.R["value"]) - np.log(100))


def test_update_lfp_semantics():
    m = _model_1d()
    new = np.zeros((24, 40))
    m.update_lfp(new, m.t)
    assert m.lfp is new                                     # 1-D keeps the array as given (gpcsd1d.py:111)
    m2 = _model_2d()
    m2.update_lfp(np.zeros((48, 30)), m2.t)
    assert m2.lfp.shape == (48, 30, 1)                      # 2-D re-applies atleast_3d (gpcsd2d.py:134)


def test_constructor_rng_coupling_matches_reference():
    """Same np.random.seed => same prior-sampled initial values and bounds as the reference constructors
    (SURVEY.md 9.8); run in a subprocess because both packages answer to different import names."""
    from oracle import ref_shim
    if not ref_shim.reference_available():
        pytest.skip("reference tree not present (GPU box)")
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r)
from oracle.ref_shim import import_reference
g = import_reference()
from gpcsd_b200.gpcsd1d import GPCSD1D
from gpcsd_b200.gpcsd2d import GPCSD2D
x = np.linspace(0, 2300, 24)[:, None]; t = np.linspace(0, 50, 40)[:, None]; lfp = np.zeros((24, 40, 2))
np.random.seed(5); a = g.gpcsd1d.GPCSD1D(lfp, x, t)
np.random.seed(5); b = GPCSD1D(lfp, x, t)
pa, pb = a.extract_model_params(), b.extract_model_params()
for k in pa: assert np.allclose(pa[k], pb[k], rtol=0, atol=0), k
for d1, d2 in ((a.R, b.R), (a.sig2n, b.sig2n), (a.spatial_cov.params['ell'], b.spatial_cov.params['ell'])):
    assert d1['min'] == d2['min'] and d1['max'] == d2['max']
    assert str(d1['prior']) == str(d2['prior'])
X = g.utility_functions.expand_grid(np.linspace(0,48,4)[:,None], np.linspace(0,220,12)[:,None]); t2 = np.arange(30.)[:,None]; l2 = np.zeros((48,30,2))
np.random.seed(6); a2 = g.gpcsd2d.GPCSD2D(l2, X, t2, ngl1=5, ngl2=9)
np.random.seed(6); b2 = GPCSD2D(l2, X, t2, ngl1=5, ngl2=9)
pa, pb = a2.extract_model_params(), b2.extract_model_params()
for k in pa: assert np.allclose(pa[k], pb[k], rtol=0, atol=0), k
for key in ('ell1', 'ell2'):
    assert a2.spatial_cov.params[key]['min'] == b2.spatial_cov.params[key]['min']
    assert a2.spatial_cov.params[key]['max'] == b2.spatial_cov.params[key]['max']
assert np.array_equal(a2.spatial_cov.gl_x_grid, b2.spatial_cov.gl_x_grid) and np.array_equal(a2.spatial_cov.gl_w_prod, b2.spatial_cov.gl_w_prod)
assert a2.R['min'] == b2.R['min'] and a2.R['max'] == b2.R['max'] and a2.eps == b2.eps
print('ok')
''' % ROOT
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-3000:]


def test_dropin_alias_package():
    code = "import sys; sys.path.insert(0, %r); sys.path.insert(0, %r); from gpcsd.gpcsd1d import GPCSD1D; from gpcsd.covariances import *; " \
           "from gpcsd.gpcsd2d import GPCSD2D; import gpcsd.predict_csd, gpcsd.utility_functions; " \
           "assert callable(fwd_model_1d) and callable(b_fwd_2d) and GPCSDInvGammaPrior; print('ok')" % (ROOT, os.path.join(ROOT, "dropin"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


def test_collective_order_rotation():
    """parallel.CollectiveOrder: concurrent members enqueue in the fixed rotation 0,1,0,1,...; outside a started phase
    turn() never waits (a single thread driving the members one after the other must not dead-lock)."""
    import threading
    from gpcsd_b200.parallel import CollectiveOrder
    order = CollectiveOrder(2)
    log = []
    for idx in (0, 0, 1, 1):                      # not started: free
        with order.turn(idx):
            log.append(idx)
    assert log == [0, 0, 1, 1]
    log.clear()
    order.start()

    def member(idx):
        for _ in range(50):
            with order.turn(idx):
                log.append(idx)
    ths = [threading.Thread(target=member, args=(i,)) for i in (1, 0)]
    [t.start() for t in ths]
    [t.join(timeout=20) for t in ths]
    order.stop()
    assert log == [0, 1] * 50


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) prints one JSON line with the contract's keys."""
    import json
    import subprocess
    import sys
    from conftest import ROOT
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "gpcsd_loglik_grad_evals_per_s" and line["unit"] == "evals/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["dtype"] == "f64"
    assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"]
