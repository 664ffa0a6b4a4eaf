"""CPU: the lock-step batched projected L-BFGS driver (gpcsd_b200/batched_opt.py) against scipy's L-BFGS-B -- the optimiser
the reference's fit() runs per restart (gpcsd1d.py:211) -- on the oracle's GPCSD objective and on textbook problems."""
import numpy as np
import scipy.optimize

from gpcsd_b200.batched_opt import batched_lbfgsb


def _rosen_batch(X, idx):
    f = np.array([scipy.optimize.rosen(x) for x in X])
    g = np.array([scipy.optimize.rosen_der(x) for x in X])
    return f, g


def test_rosenbrock_batch_matches_scipy():
    rng = np.random.default_rng(0)
    X0 = rng.uniform(-1.5, 1.5, (7, 4))
    res = batched_lbfgsb(_rosen_batch, X0, maxiter=500, gtol=1e-8, ftol=1e-15)
    assert np.all(res["fun"] < 1e-10), res
    assert np.allclose(res["x"], 1.0, atol=1e-4)


def test_bounds_are_respected_and_active_bounds_found():
    # quadratic with the unconstrained minimum outside the box: the solution sits on the bound
    c = np.array([2.0, -3.0, 0.5])

    def fun(X, idx):
        return 0.5 * np.sum((X - c) ** 2, axis=1), X - c
    bounds = [(-1.0, 1.0), (-1.0, 1.0), (None, None)]
    res = batched_lbfgsb(fun, np.zeros((3, 3)) + np.array([[0.0], [0.3], [-0.7]]), bounds=bounds)
    assert np.allclose(res["x"], np.array([1.0, -1.0, 0.5])[None, :], atol=1e-6)
    ref = scipy.optimize.minimize(lambda x: (0.5 * np.sum((x - c) ** 2), x - c), np.zeros(3), jac=True, method="L-BFGS-B",
                                  bounds=[(-1, 1), (-1, 1), (None, None)])
    assert np.allclose(res["x"][0], ref.x, atol=1e-6)


def test_infinite_bounds_like_the_reference():
    """The reference passes log(0) = -inf and log(inf) bounds (covariances.py:289, gpcsd1d.py:146)."""
    def fun(X, idx):
        return np.sum(np.cosh(X - 0.3), axis=1), np.sinh(X - 0.3)
    res = batched_lbfgsb(fun, np.array([[2.0, -1.0]]), bounds=[(-np.inf, np.inf), (-np.inf, 5.0)])
    assert np.allclose(res["x"], 0.3, atol=1e-5)


def test_gpcsd_objective_multistart_matches_scipy_lbfgsb():
    """Four prior-sampled restarts of the oracle's negative log posterior (gpcsd1d.py:153-191) optimised in lock step reach
    the same optima (objective value) as scipy's L-BFGS-B run per restart with the reference's options."""
    from oracle import gpcsd_oracle as O, synth
    x, t = synth.geometry_1d(12, 24)
    om = synth.model_1d(x, t, sig2n=1e-2)
    lfp = synth.matched_lfp(om, 12, 4)
    pri = synth.default_priors(om)
    tp_true = O.pack_tparams(om)
    rng = np.random.default_rng(5)
    starts = tp_true[None, :] + 0.4 * rng.standard_normal((4, tp_true.size))
    bounds = [(-3.0, 3.0), (-3.0, 3.0), (np.log(1.0), np.log(30.0)), (-np.inf, np.inf), (np.log(0.5), np.log(30.0)),
              (-np.inf, np.inf), (np.log(1e-8), np.log(0.5))]

    def one(tp):
        return O.obj_and_grad(om, lfp, tp, pri)

    def batch(X, idx):
        vals = [one(xr) for xr in X]
        return np.array([v[0] for v in vals]), np.array([v[1] for v in vals])
    opts = {'maxiter': 200, 'gtol': 1e-5, 'ftol': 1e7 * np.finfo(float).eps}
    res = batched_lbfgsb(batch, starts, bounds=bounds, maxiter=200, gtol=1e-5, ftol=opts['ftol'])
    for b in range(4):
        ref = scipy.optimize.minimize(one, starts[b], jac=True, method="L-BFGS-B", bounds=bounds, options=opts)
        # both stop on the same relative-decrease / projected-gradient tests, so the optima agree to that tolerance
        assert abs(res["fun"][b] - ref.fun) <= 2e-6 * max(abs(ref.fun), 1.0), (b, res["fun"][b], ref.fun, res["status"][b])
    assert res["nfev"] < 4 * 200
