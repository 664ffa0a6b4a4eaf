"""Traditional second-difference CSD estimators (comparison baseline, not part of the GPCSD hot path) --
API mirror of ``gpcsd.predict_csd`` (predict_csd.py:3-31).  Plain numpy stencils."""
import numpy as np


def predictcsd_trad_1d(lfp):
    """-(lfp[x+1] + lfp[x-1] - 2 lfp[x]) on interior contacts, zero at the two ends; lfp (nx, nt, ntrial)."""
    lfp = np.asarray(lfp)
    csd = np.zeros(lfp.shape)
    csd[1:-1] = lfp[2:] + lfp[:-2] - 2.0 * lfp[1:-1]
    return -csd


def predictcsd_trad_2d(lfp):
    """Column-wise second difference along axis 1; NaN on the two border columns; lfp (nx1, nx2, nt, ntrial)."""
    lfp = np.asarray(lfp)
    csd = np.full(lfp.shape, np.nan)
    csd[:, 1:-1] = lfp[:, 2:] + lfp[:, :-2] - 2.0 * lfp[:, 1:-1]
    return -csd
