"""Hyperpriors of the GPCSD model -- API mirror of the reference's ``gpcsd.priors`` (priors.py:14-54).

Closed-form scalar densities evaluated on the host; their values and derivatives are added to the
objective / gradient that the CUDA engine returns (gpcsd1d.py:177-186).
"""
import numpy as np
from scipy.stats import halfnorm, invgamma


class GPCSDPrior:
    def __init__(self):
        pass


class GPCSDInvGammaPrior(GPCSDPrior):
    """Inverse-gamma prior; ``lpdf`` is un-normalised exactly like priors.py:23-28."""

    def __init__(self, alpha=1, beta=1):
        super().__init__()
        self.alpha = alpha
        self.beta = beta

    def __str__(self):
        return "InvGamma(%0.2f, %0.2f)" % (self.alpha, self.beta)

    def lpdf(self, x):
        if x <= 0:
            return -np.inf
        return -(self.alpha + 1.0) * np.log(x) - self.beta / x

    def dlpdf(self, x):
        """d lpdf / dx (what autograd derives from priors.py:27)."""
        return -(self.alpha + 1.0) / x + self.beta / (x * x)

    def set_params(self, l, u):
        """Moment match so that most mass lies in [l, u] (priors.py:30-32)."""
        self.alpha = 2.0 + 9.0 * np.square((l + u) / (u - l))
        self.beta = 0.5 * (self.alpha - 1.0) * (l + u)

    def sample(self):
        return invgamma.rvs(self.alpha, scale=self.beta)


class GPCSDHalfNormalPrior(GPCSDPrior):
    """Half-normal prior; un-normalised lpdf (priors.py:46-51)."""

    def __init__(self, sd=1):
        super().__init__()
        self.sd = sd

    def __str__(self):
        return "HalfNormal(%0.2f)" % (self.sd)

    def lpdf(self, x):
        if x <= 0:
            return -np.inf
        return -0.5 * np.square(x / self.sd)

    def dlpdf(self, x):
        return -x / (self.sd * self.sd)

    def sample(self):
        return halfnorm.rvs(scale=self.sd)
