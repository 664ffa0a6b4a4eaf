"""GPCSD1D -- 1-D (laminar probe) GPCSD model on the B200 engine; drop-in for ``gpcsd.gpcsd1d.GPCSD1D``
(gpcsd1d.py:19-309)."""
import numpy as np

from ._model import GPCSDModelBase
from .covariances import *  # noqa: F401,F403  (the reference module re-exports these names)
from .covariances import GPCSD1DSpatialCovSE, GPCSDTemporalCovMatern, GPCSDTemporalCovSE
from .forward_models import *  # noqa: F401,F403
from .priors import *  # noqa: F401,F403
from .priors import GPCSDHalfNormalPrior, GPCSDInvGammaPrior
from .utility_functions import *  # noqa: F401,F403

JITTER = 1e-8  # gpcsd1d.py:17


class GPCSD1D(GPCSDModelBase):
    DIM = 1
    JITTER = JITTER
    SPATIAL_ELL_KEYS = ('ell',)

    def __init__(self, lfp, x, t, a=None, b=None, ngl=100, spatial_cov=None, temporal_cov_list=None, R_prior=None,
                 sig2n_prior=None, distributed=False, distributed_restarts=False):
        """
        :param lfp: LFP array (n_spatial, n_time, n_trials); rescale to roughly unit standard deviation
        :param x: electrode positions (n_spatial, 1), microns
        :param t: time points (n_time, 1), milliseconds
        :param a, b: integration limits of the forward model (default: min / max of x)
        :param ngl: Gauss-Legendre order
        :param spatial_cov: GPCSD1DSpatialCovSE instance (default constructed from x, a, b, ngl)
        :param temporal_cov_list: list of temporal covariance objects (default [SE, Matern])
        :param R_prior: prior on the cylinder radius R (default inverse-gamma matched to the probe)
        :param sig2n_prior: prior on the noise variance, or a list with one prior per electrode
        :param distributed: True (or a torch.distributed group) shards the trials over the ranks
        :param distributed_restarts: True (or a group) shards fit()'s multi-start restarts over the ranks instead
        """
        self.lfp = np.atleast_3d(lfp)
        self.x = x
        self.t = t
        self.a = np.min(x) if a is None else a
        self.b = np.max(x) if b is None else b
        self.ngl = ngl
        self._group = distributed if distributed else None
        self._restart_group = distributed_restarts if distributed_restarts else None
        if spatial_cov is None:
            spatial_cov = GPCSD1DSpatialCovSE(x, a=self.a, b=self.b, ngl=ngl)
        self.spatial_cov = spatial_cov
        if temporal_cov_list is None:
            temporal_cov_list = [GPCSDTemporalCovSE(t), GPCSDTemporalCovMatern(t)]
        self.temporal_cov_list = temporal_cov_list
        xs = self.x.squeeze()
        dmin, span = np.min(np.diff(xs)), np.max(xs) - np.min(xs)
        if R_prior is None:
            R_prior = GPCSDInvGammaPrior()
            R_prior.set_params(dmin, 0.5 * span)
        self.R = {'value': R_prior.sample(), 'prior': R_prior, 'min': 0.5 * dmin,
                  'max': 0.8 * (np.max(self.x) - np.min(self.x))}
        if sig2n_prior is None:
            sig2n_prior = GPCSDHalfNormalPrior(0.1)
        if isinstance(sig2n_prior, list):
            n = len(sig2n_prior)
            self.sig2n = {'value': np.array([p.sample() for p in sig2n_prior]), 'prior': sig2n_prior,
                          'min': [1e-8] * n, 'max': [0.5] * n}
        else:
            self.sig2n = {'value': sig2n_prior.sample(), 'prior': sig2n_prior, 'min': 1e-8, 'max': 0.5}

    def _quadrature(self):
        return dict(gl_x=self.spatial_cov.gl_x, gl_w=self.spatial_cov.gl_w)

    def __str__(self):
        s = "GPCSD1D object\n"
        s += "LFP shape: (%d, %d, %d)\n" % self.lfp.shape[:3]
        s += "Integration bounds: (%d, %d)\n" % (self.a, self.b)
        s += "Integration number points: %d\n" % self.ngl
        s += "R parameter prior: %s\n" % str(self.R['prior'])
        s += "R parameter value %0.4g\n" % self.R['value']
        s += "Spatial covariance ell prior: %s\n" % str(self.spatial_cov.params['ell']['prior'])
        s += "Spatial covariance ell value %0.4g\n" % self.spatial_cov.params['ell']['value']
        return s + self._str_temporal()

    def extract_model_params(self):
        ells, s2 = self._temporal_lists()
        return {'R': self.R['value'], 'sig2n': self.sig2n['value'],
                'spatial_ell': self.spatial_cov.params['ell']['value'],
                'temporal_ell_list': ells, 'temporal_sigma2_list': s2}

    def restore_model_params(self, params):
        self.R['value'] = params['R']
        self.sig2n['value'] = params['sig2n']
        self.spatial_cov.params['ell']['value'] = params['spatial_ell']
        self._restore_temporal(params)

    def update_lfp(self, new_lfp, t, x=None):
        """Swap the data (and optionally the geometry) keeping the hyperparameters (gpcsd1d.py:104-111;
        like the reference, new_lfp is stored as given -- pass a 3-D array)."""
        if x is not None:
            self.x = x
            self.spatial_cov.x = x
        if t is not self.t:
            self.t = t
            for tcov in self.temporal_cov_list:
                tcov.t = t
        self.lfp = new_lfp
        self._invalidate_lfp()

    def fit(self, n_restarts=10, method='L-BFGS-B', fix_R=False, verbose=False,
            options={'maxiter': 1000, 'disp': False, 'gtol': 1e-5, 'ftol': 1e7 * np.finfo(float).eps}, n_workers=2, lockstep=None):
        """MAP fit by multi-start bounded L-BFGS-B in log space (gpcsd1d.py:130-246).  Every objective
        evaluation is one fused loglik+gradient pass on the GPU; ``n_workers`` restarts run concurrently (extension)."""
        return self._fit(n_restarts, method, fix_R, verbose, options, n_workers=n_workers, lockstep=lockstep)

    def sample_prior(self, ntrials, device=False, seed=0):
        """CSD draws from the GP prior at the electrode sites (gpcsd1d.py:295-309): Cholesky factors of the CSD kernel
        (+ JITTER) and of Kt on the device (gpcsd_cholesky), then Ls Z_r Lt^T for all trials as two DMMA GEMMs.
        device=False (reference behaviour): Z is drawn with np.random.normal in the reference's order (one (nx, nt) block per
        trial from the global numpy RNG) and a host array is returned.  device=True: Z comes from the device's Philox4x32-10
        generator (`seed`), nothing touches the host, and a CUDA tensor (nx, nt, ntrials) is returned."""
        from . import devops
        nt, nx = self.t.shape[0], self.x.shape[0]
        Lt = devops.cholesky_device(self._kt_total())
        Ls = devops.cholesky_device(self.spatial_cov.compute_Ks() + JITTER * np.eye(nx))
        if device:
            out, n = devops.sample_gp_device(Ls, Lt, ntrials, seed)
            return out[:, :, :n]
        rand = np.stack([np.random.normal(0, 1, (nx, nt)) for _ in range(ntrials)], axis=2)
        out, n = devops.sample_gp_device(Ls, Lt, ntrials, 0, rand=rand)
        return np.ascontiguousarray(out[:, :, :n].cpu().numpy())
