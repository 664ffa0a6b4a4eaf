"""Grid / Kronecker helpers -- API mirror of ``gpcsd.utility_functions`` (utility_functions.py:7-64).

``normalize``, ``sort_grid``, ``expand_grid``, ``reduce_grid`` are trivial host-side index helpers.
``comp_eig_D`` runs on the GPU (cuSOLVER syevd through the C ABI) and ``mykron`` is kept for API
compatibility only -- the engine never materialises a Kronecker product.
"""
import numpy as np


def normalize(x):
    """Scale each trial to unit max-abs (utility_functions.py:7-8)."""
    return x / np.max(np.abs(x), axis=(0, 1))


def sort_grid(x):
    """Sort (n, 2) points by first column, ties by second (utility_functions.py:10-13)."""
    x = np.asarray(x)
    return x[np.lexsort((x[:, 1], x[:, 0]))]


def expand_grid(x1, x2):
    """All pairs (a, b), a in x1 (slow index), b in x2 (fast index) -> (len(x1)*len(x2), 2)
    (utility_functions.py:15-23)."""
    a = np.asarray(x1, dtype=np.float64).reshape(-1)
    b = np.asarray(x2, dtype=np.float64).reshape(-1)
    return np.squeeze(np.stack([np.repeat(a, b.size), np.tile(b, a.size)], axis=1))


def reduce_grid(x):
    """Unique sorted coordinates of each column (utility_functions.py:25-33)."""
    return np.unique(x[:, 0]), np.unique(x[:, 1])


def mykron(A, B):
    """Dense Kronecker product with the reference's block order (utility_functions.py:35-42)."""
    A, B = np.asarray(A), np.asarray(B)
    return np.einsum("ij,kl->ikjl", A, B).reshape(A.shape[0] * B.shape[0], A.shape[1] * B.shape[1])


def comp_eig_D(Ks, Kt, sig2n):
    """Eigenvectors of Ks and Kt and the Kronecker eigenvalue vector
    Dvec[i*nt + j] = ls_i * lt_j + sig2n (scalar) or + sig2n[i] (vector, i = ascending spatial
    eigenvalue index) -- utility_functions.py:44-64.  Runs on the GPU; returns host arrays."""
    from . import devops
    return devops.comp_eig_D(Ks, Kt, sig2n)
