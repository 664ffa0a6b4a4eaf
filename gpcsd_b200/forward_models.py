"""Forward-model weight functions and CSD->LFP trapezoid operators -- API mirror of
``gpcsd.forward_models`` (forward_models.py:9-81).

``b_fwd_1d`` / ``b_fwd_2d`` are the weight functions the covariance kernels fuse on the GPU
(gpcsd_fwd_weights_{1d,2d}); the host versions here serve callers that use them directly.
``fwd_model_1d`` / ``fwd_model_2d`` (synthetic-data generators) run on the device as ONE
weight-matrix kernel + ONE DMMA GEMM each instead of the reference's Python double loop over time x location.
"""
import numpy as np


def b_fwd_1d(r, R):
    """sqrt((r/R)^2 + 1) - sqrt((r/R)^2)  (forward_models.py:9-17)."""
    q = np.square(np.divide(r, R))
    return np.sqrt(q + 1) - np.sqrt(q)


def fwd_model_1d(arr, x, z, R, varsigma=1):
    """LFP at z from CSD ``arr`` (nx, nt) sampled at x: R/(2 varsigma) * trapz_x b(z_i - x) arr[:, t]
    (forward_models.py:20-39).  On the device: gpcsd_fwd_operator_1d builds the (nz, nx) weight matrix (b_fwd_1d x
    trapezoid weights x R/(2 varsigma)) and one gpcsd_dgemm applies it to all time points."""
    from . import devops
    return devops.fwd_apply_1d(arr, x, z, R, varsigma)


def b_fwd_2d(delta1, delta2, R, eps, w=None):
    """log(R+eps+sqrt((R+eps)^2+w^2)) - log(eps+sqrt(eps^2+w^2)), w = |delta| (forward_models.py:42-54)."""
    if w is None:
        w = np.array(np.sqrt(np.square(delta1) + np.square(delta2)))
    return np.log(R + eps + np.sqrt((R + eps) ** 2 + w ** 2)) - np.log(eps + np.sqrt(eps ** 2 + w ** 2))


def fwd_model_2d(arr, x1, x2, z, R, eps, varsigma=1):
    """LFP at z (nz, 2) from CSD ``arr`` (nx1, nx2, nt) on the grid x1 x x2: double trapezoid of
    b_fwd_2d * arr (forward_models.py:57-81; the 1/(4 pi varsigma) factor is omitted there too).  On the device:
    gpcsd_fwd_operator_2d + one gpcsd_dgemm."""
    from . import devops
    arr = np.asarray(arr, dtype=np.float64)
    if arr.ndim == 4 and arr.shape[3] == 1:       # callers pass (nx1, nx2, nt, 1) (sim_from_gp_2D.py:70); the reference squeezes it
        arr = arr[..., 0]
    return devops.fwd_apply_2d(arr, x1, x2, z, R, eps)
