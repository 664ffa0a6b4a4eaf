"""TEST INFRASTRUCTURE ONLY -- import the untouched reference in THIS container.

The reference (``/root/reference/src/gpcsd``) imports ``autograd.numpy`` (HIPS autograd, un-pinned in
the reference's ``setup.py:30``; not installed here, no network) and calls ``scipy.integrate.trapz``
(removed from scipy >= 1.14).  This shim makes the *forward* code importable without modifying it:

* ``autograd.numpy`` -> plain ``numpy`` (every ``np.*`` the reference calls exists in numpy),
* ``autograd.grad``  -> raises (the gradient path cannot run: gradient parity is therefore UNPINNED
  by the reference itself; see oracle/gpcsd_oracle.py header),
* ``scipy.integrate.trapz`` -> ``scipy.integrate.trapezoid``.

``/root/reference`` does not exist on the GPU box, so this module is only used by
``oracle/make_golden.py`` (fixture generator, run here) and by CPU tests that skip when the
reference tree is absent.
"""
import os
import sys
import types

REFERENCE_SRC = "/root/reference/src"


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_SRC, "gpcsd"))


def import_reference():
    """Return the reference's modules as a namespace: .gpcsd1d, .gpcsd2d, .covariances, ..."""
    if not reference_available():
        raise ImportError("reference tree not present at %s" % REFERENCE_SRC)
    if "gpcsd" in sys.modules and not getattr(sys.modules["gpcsd"], "__file__", "").startswith(REFERENCE_SRC):
        raise ImportError("a different package named 'gpcsd' is already imported; run the shim in a fresh process")
    import numpy
    import scipy
    import scipy.integrate
    import scipy.optimize
    import scipy.special
    import scipy.stats

    if "autograd" not in sys.modules:
        ag = types.ModuleType("autograd")

        def _no_grad(*a, **k):
            raise RuntimeError("HIPS autograd is not installed: the reference gradient path cannot run here")

        ag.grad = _no_grad
        ag.numpy = numpy
        sys.modules["autograd"] = ag
        sys.modules["autograd.numpy"] = numpy
    if not hasattr(scipy.integrate, "trapz"):
        scipy.integrate.trapz = scipy.integrate.trapezoid
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    import gpcsd.covariances
    import gpcsd.forward_models
    import gpcsd.gpcsd1d
    import gpcsd.gpcsd2d
    import gpcsd.predict_csd
    import gpcsd.priors
    import gpcsd.utility_functions

    return sys.modules["gpcsd"]
