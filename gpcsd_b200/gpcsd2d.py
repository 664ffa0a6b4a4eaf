"""GPCSD2D -- 2-D (Neuropixels-style) GPCSD model on the B200 engine; drop-in for
``gpcsd.gpcsd2d.GPCSD2D`` (gpcsd2d.py:18-360)."""
import numpy as np

from ._model import GPCSDModelBase
from .covariances import GPCSD2DSpatialCovSE, GPCSDTemporalCovMatern, GPCSDTemporalCovSE
from .priors import GPCSDHalfNormalPrior, GPCSDInvGammaPrior
from .utility_functions import comp_eig_D, mykron, reduce_grid  # noqa: F401

JITTER = 1e-7  # gpcsd2d.py:16


class GPCSD2D(GPCSDModelBase):
    DIM = 2
    JITTER = JITTER
    SPATIAL_ELL_KEYS = ('ell1', 'ell2')

    def __init__(self, lfp, x, t, a1=None, b1=None, a2=None, b2=None, ngl1=20, ngl2=60, spatial_cov=None,
                 temporal_cov_list=None, R_prior=None, sig2n_prior=None, eps=None, distributed=False, distributed_restarts=False):
        """
        :param lfp: LFP array (n_spatial, n_time, n_trials); rescale to roughly unit standard deviation
        :param x: electrode positions (n_spatial, 2), microns
        :param t: time points (n_time, 1), milliseconds
        :param a1, b1, a2, b2: integration limits per spatial dimension (default: min / max of x columns)
        :param ngl1, ngl2: Gauss-Legendre orders
        :param eps: zero-charge offset in front of the array (default 5 * smallest electrode spacing)
        :param distributed: True (or a torch.distributed group) shards the trials over the ranks
        :param distributed_restarts: True (or a group) shards fit()'s multi-start restarts over the ranks instead
        """
        self.lfp = np.atleast_3d(lfp)
        self.x = x
        self.t = t
        self.a1 = np.min(x[:, 0]) if a1 is None else a1
        self.b1 = np.max(x[:, 0]) if b1 is None else b1
        self.a2 = np.min(x[:, 1]) if a2 is None else a2
        self.b2 = np.max(x[:, 1]) if b2 is None else b2
        self.ngl1, self.ngl2 = ngl1, ngl2
        self._group = distributed if distributed else None
        self._restart_group = distributed_restarts if distributed_restarts else None
        if spatial_cov is None:
            spatial_cov = GPCSD2DSpatialCovSE(self.x, a1=self.a1, b1=self.b1, a2=self.a2, b2=self.b2, ngl1=ngl1, ngl2=ngl2)
        self.spatial_cov = spatial_cov
        if temporal_cov_list is None:
            temporal_cov_list = [GPCSDTemporalCovSE(t), GPCSDTemporalCovMatern(t)]
        self.temporal_cov_list = temporal_cov_list
        x1, x2 = reduce_grid(x)
        min_delta_x = np.min([np.min(np.diff(x1)), np.min(np.diff(x2))])
        max_delta_x = np.max([self.b1 - self.a1, self.b2 - self.a2])
        if R_prior is None:
            R_prior = GPCSDInvGammaPrior()
            R_prior.set_params(min_delta_x, 0.5 * max_delta_x)
        self.R = {'value': R_prior.sample(), 'prior': R_prior, 'min': 0.5 * min_delta_x, 'max': 0.8 * max_delta_x}
        self.eps = 5 * min_delta_x if eps is None else eps
        if sig2n_prior is None:
            sig2n_prior = GPCSDHalfNormalPrior(1.0)
        if isinstance(sig2n_prior, list):
            n = len(sig2n_prior)
            self.sig2n = {'value': np.array([p.sample() for p in sig2n_prior]), 'prior': sig2n_prior,
                          'min': [1e-8] * n, 'max': [10.0] * n}
        else:
            self.sig2n = {'value': sig2n_prior.sample(), 'prior': sig2n_prior, 'min': 1e-8, 'max': 10.0}

    def _quadrature(self):
        sc = self.spatial_cov
        return dict(gl_x1=sc.gl_x1, gl_w1=sc.gl_w1, gl_x2=sc.gl_x2, gl_w2=sc.gl_w2)

    def __str__(self):
        s = "GPCSD1D object\n"  # sic: the reference's 2-D header says 1D (gpcsd2d.py:82)
        s += "LFP shape: (%d, %d, %d)\n" % self.lfp.shape[:3]
        s += "Integration bounds: (%d, %d), (%d, %d)\n" % (self.a1, self.b1, self.a2, self.b2)
        s += "Integration number points: %d, %d\n" % (self.ngl1, self.ngl2)
        s += "R parameter prior: %s\n" % str(self.R['prior'])
        s += "R parameter value %0.4g\n" % self.R['value']
        for k, d in ((1, 'ell1'), (2, 'ell2')):
            s += "Spatial covariance ell prior (dim %d): %s\n" % (k, str(self.spatial_cov.params[d]['prior']))
            s += "Spatial covariance ell value (dim %d) %0.4g\n" % (k, self.spatial_cov.params[d]['value'])
        return s + self._str_temporal()

    def extract_model_params(self):
        ells, s2 = self._temporal_lists()
        return {'R': self.R['value'], 'eps': self.eps, 'sig2n': self.sig2n['value'],
                'spatial_ell1': self.spatial_cov.params['ell1']['value'],
                'spatial_ell2': self.spatial_cov.params['ell2']['value'],
                'temporal_ell_list': ells, 'temporal_sigma2_list': s2}

    def restore_model_params(self, params):
        self.R['value'] = params['R']
        self.eps = params['eps']
        self.sig2n['value'] = params['sig2n']
        self.spatial_cov.params['ell1']['value'] = params['spatial_ell1']
        self.spatial_cov.params['ell2']['value'] = params['spatial_ell2']
        self._restore_temporal(params)

    def update_lfp(self, new_lfp, t, x=None):
        """gpcsd2d.py:127-134 (the 2-D class re-applies atleast_3d and resets the spatial geometry)."""
        if x is not None:
            self.x = x
            self.spatial_cov.reset_x(x)
        if t is not self.t:
            self.t = t
            for tcov in self.temporal_cov_list:
                tcov.t = t
        self.lfp = np.atleast_3d(new_lfp)
        self._invalidate_lfp()

    def fit(self, n_restarts=10, method='L-BFGS-B', fix_R=False, verbose=False, profile=False,
            options={'maxiter': 500, 'disp': False, 'gtol': 1e-5, 'ftol': 1e7 * np.finfo(float).eps}, n_workers=2, lockstep=None):
        """MAP fit by multi-start bounded L-BFGS-B in log space (gpcsd2d.py:153-287).  ``profile=True`` keeps
        the reference's hook: profile one objective+gradient evaluation from a prior-sampled start with
        cProfile (files objfunstats / gradobjfunstats) and return."""
        if profile:
            import cProfile
            tparams0 = self._sample_tparams0(fix_R)
            cProfile.runctx('self.obj_fun(tparams0, fix_R)', None, locals(), filename='objfunstats')
            cProfile.runctx('self.obj_fun_and_grad(tparams0, fix_R)', None, locals(), filename='gradobjfunstats')
            return
        return self._fit(n_restarts, method, fix_R, verbose, options, n_workers=n_workers, lockstep=lockstep)

    def sample_prior(self, ntrials, type="csd", seed=1, device=False):
        """CSD and/or LFP draws from the GP prior (gpcsd2d.py:336-360); returns (csd, lfp), NaN-filled when not
        requested.  Cholesky factors on the device (gpcsd_cholesky), Ls Z_r Lt^T as two DMMA GEMMs.  device=False: the
        reference's np.random.seed(seed) + np.random.normal draw, host arrays.  device=True: Philox4x32-10 normals on the
        device (`seed`), CUDA tensors (None for the kind not requested)."""
        from . import devops
        nt, nx = self.t.shape[0], self.x.shape[0]
        want_csd, want_lfp = type in ("csd", "both"), type in ("lfp", "both")
        Ls_csd = Ls_lfp = None
        if want_csd:
            Ls_csd = devops.cholesky_device(self.spatial_cov.compute_Ks() + JITTER * np.eye(nx))
        if want_lfp:
            Ls_lfp = devops.cholesky_device(self.spatial_cov.compKphi_2d(R=self.R['value'], eps=self.eps) + JITTER * np.eye(nx))
        Lt = devops.cholesky_device(self._kt_total())
        if device:
            outs = []
            for Ls in (Ls_csd, Ls_lfp):
                if Ls is None:
                    outs.append(None)
                else:
                    o, n = devops.sample_gp_device(Ls, Lt, ntrials, seed)      # same Z for both kinds, as in the reference
                    outs.append(o[:, :, :n])
            return tuple(outs)
        np.random.seed(seed)
        csd = np.nan * np.zeros((nx, nt, ntrials))
        lfp = np.nan * np.zeros((nx, nt, ntrials))
        rand_samp = np.random.normal(0, 1, (nx, nt, ntrials))
        if want_csd:
            o, n = devops.sample_gp_device(Ls_csd, Lt, ntrials, 0, rand=rand_samp)
            csd = np.ascontiguousarray(o[:, :, :n].cpu().numpy())
        if want_lfp:
            o, n = devops.sample_gp_device(Ls_lfp, Lt, ntrials, 0, rand=rand_samp)
            lfp = np.ascontiguousarray(o[:, :, :n].cpu().numpy())
        return csd, lfp
