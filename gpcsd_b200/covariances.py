"""Spatial and temporal covariance objects -- API mirror of ``gpcsd.covariances`` (covariances.py:12-305).

Same class names, constructor arguments, ``params`` dictionaries ({'value','prior','min','max'}) and
method names as the reference, so scripts that mutate ``params[...]['value']`` keep working.  The
matrices themselves are built by the CUDA kernels of libgpcsd_b200.so (see devops.py); inside
``loglik``/``predict`` the engine fuses these builders and never round-trips through the host.
"""
import numpy as np
import scipy
import scipy.special

from . import devops
from .forward_models import *  # noqa: F401,F403  (callers rely on `from gpcsd.covariances import *` re-exports)
from .priors import *  # noqa: F401,F403
from .priors import GPCSDHalfNormalPrior, GPCSDInvGammaPrior
from .utility_functions import expand_grid, reduce_grid


def _gl_on_interval(a, b, n):
    """Gauss-Legendre nodes / weights mapped to [a, b] (covariances.py:22-27)."""
    u, w = scipy.special.roots_legendre(n)
    return 0.5 * (u + 1.0) * (b - a) + a, 0.5 * (b - a) * w


def _default_ell_prior(lo, hi):
    p = GPCSDInvGammaPrior()
    p.set_params(lo, hi)
    return p


class GPCSD1DSpatialCov:
    """Geometry + quadrature of the 1-D (laminar probe) forward model (covariances.py:12-27)."""

    def __init__(self, x, a, b, ngl):
        self.x = x
        self.a = np.min(x) if a is None else a
        self.b = np.max(x) if b is None else b
        self.ngl = ngl
        self.gl_x, self.gl_w = _gl_on_interval(self.a, self.b, ngl)


class GPCSD1DSpatialCovSE(GPCSD1DSpatialCov):
    """Squared-exponential CSD kernel pushed through the 1-D forward model (covariances.py:29-96)."""

    def __init__(self, x, ell_prior=None, a=None, b=None, ngl=100):
        super().__init__(x, a, b, ngl)
        xs = self.x.squeeze()
        dmin, span = np.min(np.diff(xs)), np.max(xs) - np.min(xs)
        if ell_prior is None:
            ell_prior = _default_ell_prior(1.2 * dmin, 0.8 * span)
        self.params = {'ell': {'value': ell_prior.sample(), 'prior': ell_prior, 'min': 0.5 * dmin, 'max': span}}

    def compute_Ks(self):
        """CSD-CSD correlation at the electrode sites (covariances.py:50-56)."""
        return devops.se_matrix(self.x, self.x, self.params['ell']['value'])

    def compKphig_1d(self, z, R):
        """LFP(x) - CSD(z) cross-covariance, (nx, nz) (covariances.py:58-72)."""
        return devops.kphig_1d(self.x, self.gl_x, self.gl_w, z, R, self.params['ell']['value'])

    def compKphi_1d(self, R, xp=None):
        """LFP(x) - LFP(xp) covariance, (nx, nxp) (covariances.py:74-96)."""
        return devops.kphi_1d(self.x, self.gl_x, self.gl_w, R, self.params['ell']['value'], xp=xp)


class GPCSD2DSpatialCov:
    """Geometry + product quadrature of the 2-D forward model (covariances.py:99-137).

    The reference caches (G x G) squared-distance matrices and an (nx x G) distance matrix on the host
    (2 x 104 MB at the Neuropixels configuration).  They are not needed here: the CSD kernel on the
    product grid is the Kronecker product of two 1-D factors and distances are recomputed in-kernel.
    ``gl_x_grid`` / ``gl_w_prod`` are kept because callers read them."""

    def __init__(self, x, a1, b1, a2, b2, ngl1, ngl2):
        self.x = x
        self.a1, self.b1, self.a2, self.b2 = a1, b1, a2, b2
        self.ngl1, self.ngl2 = ngl1, ngl2
        self.gl_x1, self.gl_w1 = _gl_on_interval(a1, b1, ngl1)
        self.gl_x2, self.gl_w2 = _gl_on_interval(a2, b2, ngl2)
        self.gl_x_grid = expand_grid(self.gl_x1, self.gl_x2)
        self.gl_w_prod = np.prod(expand_grid(self.gl_w1, self.gl_w2), axis=1, keepdims=True)
        self._quad = None

    def _device_quad(self):
        if self._quad is None:
            self._quad = devops.Quad2D(self.gl_x1, self.gl_w1, self.gl_x2, self.gl_w2)
        return self._quad

    def reset_x(self, x_new):
        """New electrode coordinates (covariances.py:133-137); nothing is cached per-x here."""
        self.x = x_new


class GPCSD2DSpatialCovSE(GPCSD2DSpatialCov):
    """Product squared-exponential CSD kernel through the 2-D forward model (covariances.py:140-232)."""

    def __init__(self, x, ell_prior1=None, ell_prior2=None, a1=None, b1=None, a2=None, b2=None, ngl1=100, ngl2=100):
        super().__init__(x, a1, b1, a2, b2, ngl1, ngl2)
        x1, x2 = reduce_grid(x)
        d1, d2 = np.min(np.diff(x1)), np.min(np.diff(x2))
        if ell_prior1 is None:
            ell_prior1 = _default_ell_prior(2.0 * d1, 2.0 * (np.max(x1) - np.min(x1)))
        if ell_prior2 is None:
            ell_prior2 = _default_ell_prior(2.0 * d2, np.max(x2) - np.min(x2))
        ell1 = ell_prior1.sample()
        ell2 = ell_prior2.sample()
        # `5.0 * max - min` (not 5 * (max - min)) is the reference's formula (covariances.py:169); kept.
        self.params = {'ell1': {'value': ell1, 'prior': ell_prior1, 'min': d1, 'max': 5.0 * np.max(x1) - np.min(x1)},
                       'ell2': {'value': ell2, 'prior': ell_prior2, 'min': d2, 'max': np.max(x2) - np.min(x2)}}

    def compute_Ks(self):
        """CSD-CSD correlation at the electrode sites (covariances.py:177-186)."""
        k1 = devops.se_matrix(self.x[:, 0], self.x[:, 0], self.params['ell1']['value'])
        k2 = devops.se_matrix(self.x[:, 1], self.x[:, 1], self.params['ell2']['value'])
        return k1 * k2

    def compKphig_2d(self, z, R, eps):
        """LFP(x) - CSD(z) cross-covariance, (nx, nz) (covariances.py:188-202)."""
        return devops.kphig_2d(self._device_quad(), self.x, z, R, eps, self.params['ell1']['value'],
                               self.params['ell2']['value'])

    def compKphi_2d(self, R, eps, xp=None):
        """LFP(x) - LFP(xp) covariance (covariances.py:204-232)."""
        return devops.kphi_2d(self._device_quad(), self.x, R, eps, self.params['ell1']['value'],
                              self.params['ell2']['value'], xp=xp)


class GPCSDTemporalCov:
    def __init__(self, t):
        self.t = t


class _TemporalKernel(GPCSDTemporalCov):
    KIND = None
    SIGMA2_MIN = 1e-8

    def __init__(self, t, ell_prior=None, sigma2_prior=None):
        super().__init__(t)
        tf = np.asarray(self.t).flatten()
        dmin, span = np.min(np.diff(tf)), np.max(tf) - np.min(tf)
        if ell_prior is None:
            ell_prior = _default_ell_prior(1.2 * dmin, 0.8 * span)
        if sigma2_prior is None:
            sigma2_prior = GPCSDHalfNormalPrior(1.0)
        ell = ell_prior.sample()
        sigma2 = sigma2_prior.sample()
        self.params = {'ell': {'value': ell, 'prior': ell_prior, 'min': 0.5 * dmin, 'max': span},
                       'sigma2': {'value': sigma2, 'prior': sigma2_prior, 'min': self.SIGMA2_MIN, 'max': np.inf}}

    def compute_Kt(self, t=None, tprime=None):
        """Temporal covariance between rows ``t`` and columns ``tprime`` (both default to self.t)."""
        t = self.t if t is None else t
        tprime = self.t if tprime is None else tprime
        return devops.kt(self.KIND, self.params['ell']['value'], self.params['sigma2']['value'], t, tprime)


class GPCSDTemporalCovSE(_TemporalKernel):
    """sigma2 * exp(-0.5 d^2 / ell^2) (covariances.py:240-271)."""
    KIND = 0
    SIGMA2_MIN = 1e-8


class GPCSDTemporalCovMatern(_TemporalKernel):
    """Matern nu=1/2: sigma2 * exp(-|d| / ell) (covariances.py:274-305); sigma2 lower bound is 0 there."""
    KIND = 1
    SIGMA2_MIN = 0
