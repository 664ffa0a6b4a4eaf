"""CPU model of the single-exchange Householder tridiagonalisation implemented by tridiag_cluster_kernel
(gpcsd_b200/csrc/gpcsd_eig.cu): per column, every participant receives p of the previous column and the owner's row as it
stood BEFORE the previous reflector was applied, finishes the previous column (w = p - tau/2 (p.v) v), rebuilds the current
row from the received one, builds the reflector with the kernel's formulas (1 + |alpha|/|beta|, sign(alpha)/(|alpha|+|beta|)),
and only then updates its own rows.  The test pins that algebra against LAPACK."""
import numpy as np
import pytest
import scipy.linalg


def tridiag_single_exchange(M):
    n = M.shape[0]
    A = M.copy()                      # rows as the owning warps hold them (updated lazily, one column behind)
    d, e, tau = np.zeros(n), np.zeros(n), np.zeros(n)
    V = np.zeros((n, n))
    vprev, tprev, p_prev = np.zeros(n), 0.0, np.zeros(n)
    rowb = A[0].copy()                # exchange 0: row 0
    for k in range(n - 1):
        # finish column k-1
        c = 0.5 * tprev * float(p_prev @ vprev)
        w = p_prev - c * vprev
        if k > 0:
            A[k:] -= np.outer(vprev[k:], w) + np.outer(w[k:], vprev)         # own rows i >= k
        if k == n - 2:
            break
        row_next = A[k + 1].copy()    # shipped with the next exchange: row k+1 BEFORE reflector k
        # row k of the current matrix rebuilt from the received pre-update row
        vk = 1.0 if k > 0 else 0.0
        wk = (p_prev[k] if k > 0 else 0.0) - c * vk
        x = rowb - vk * w - wk * vprev
        d[k] = rowb[k] - 2.0 * vk * wk
        alpha = x[k + 1]
        xnorm2 = float(x[k + 2:] @ x[k + 2:])
        v = np.zeros(n)
        v[k + 1] = 1.0
        t, beta = 0.0, alpha
        if xnorm2 > 0.0:
            s2 = alpha * alpha + xnorm2
            rn = 1.0 / np.sqrt(s2)
            ab = s2 * rn
            beta = -np.copysign(ab, alpha)
            t = 1.0 + abs(alpha) * rn
            v[k + 2:] = x[k + 2:] * np.copysign(1.0 / (abs(alpha) + ab), alpha)
        e[k], tau[k], V[k] = beta, t, v
        # symv on the own rows i > k
        p = np.zeros(n)
        p[k + 1:] = t * (A[k + 1:] @ v)
        vprev, tprev, p_prev, rowb = v, t, p, row_next
    d[n - 2], e[n - 2], d[n - 1] = A[n - 2, n - 2], A[n - 2, n - 1], A[n - 1, n - 1]
    return d, e, V, tau


@pytest.mark.parametrize("n", [3, 4, 9, 24, 65])
def test_single_exchange_tridiagonalisation_model(n):
    rng = np.random.default_rng(n)
    for trial in range(2):
        if trial == 0:
            M = rng.standard_normal((n, n))
            M = M + M.T
        else:
            tt = np.arange(n) * 0.7
            M = np.exp(-0.5 * (tt[:, None] - tt[None, :]) ** 2 / 30.0)     # degenerate tail: tiny trailing columns
        d, e, V, tau = tridiag_single_exchange(M)
        sc = np.max(np.abs(M))
        lam = scipy.linalg.eigvalsh_tridiagonal(d, e[: n - 1])
        assert np.max(np.abs(lam - np.linalg.eigvalsh(M))) < 1e-13 * sc * n
        # H = H_0 H_1 ... is orthogonal and H^T M H = T
        H = np.eye(n)
        for k in range(n - 2):
            H = H @ (np.eye(n) - tau[k] * np.outer(V[k], V[k]))
        T = np.diag(d) + np.diag(e[: n - 1], 1) + np.diag(e[: n - 1], -1)
        assert np.max(np.abs(H.T @ H - np.eye(n))) < 1e-13
        assert np.max(np.abs(H.T @ M @ H - T)) < 1e-12 * sc * n
