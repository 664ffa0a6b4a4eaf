"""CPU model of the single-exchange Householder tridiagonalisation implemented by tridiag_cluster_kernel
(gpcsd_b200/csrc/gpcsd_eig.cu).  Per column every row owner sends the pair (p_i of the column just built, a_{i,k+1} = its
element of the NEXT pivot column); every participant then finishes the previous column (w = p - tau/2 (p.v) v), rebuilds
the current pivot row from the received elements, and builds the reflector with the kernel's formulas
(1 + |alpha|/|beta|, sign(alpha)/(|alpha|+|beta|)).  The symv runs on the UNNORMALISED pivot column and on rows that have
not yet been updated by the previous reflector (A' x = A x - v (w.x) - w (v.x)); the own rows catch up after the send.
The test pins that algebra against LAPACK."""
import numpy as np
import pytest
import scipy.linalg


def tridiag_single_exchange(M):
    n = M.shape[0]
    A = M.copy()                      # rows as the owning warps hold them (updated lazily: one reflector behind at the symv)
    d, e, tau = np.zeros(n), np.zeros(n), np.zeros(n)
    V = np.zeros((n, n))
    vprev, tprev = np.zeros(n), 0.0
    p_recv, r_recv = np.zeros(n), A[:, 0].copy()      # exchange 0: column 0 of every row, p = 0
    for k in range(n - 1):
        # finish column k-1 (v^{k-1}_k = 1)
        c = 0.5 * tprev * float(p_recv[k:] @ vprev[k:])
        w = np.zeros(n)
        w[k:] = p_recv[k:] - c * vprev[k:]
        vi, wi = vprev.copy(), p_recv - c * vprev      # per own row i: v_i, w_i
        if k == n - 2:
            A[k:] -= np.outer(vi[k:], w) + np.outer(wi[k:], vprev)
            break
        wk, wk1, vk1 = p_recv[k] - c, p_recv[k + 1] - c * vprev[k + 1], vprev[k + 1]
        d[k] = r_recv[k] - 2.0 * wk
        alpha = r_recv[k + 1] - wk1 - wk * vk1
        x = np.zeros(n)
        x[k + 2:] = (r_recv[k + 2:] - w[k + 2:]) - wk * vprev[k + 2:]
        xnorm2, wx, vx = float(x @ x), float(w @ x), float(vprev @ x)
        sx = A @ x                    # rows as held: BEFORE the update by reflector k-1
        t, beta, scal = 0.0, alpha, 0.0
        if xnorm2 > 0.0:
            s2 = alpha * alpha + xnorm2
            rn = 1.0 / np.sqrt(s2)
            ab = s2 * rn
            beta = -np.copysign(ab, alpha)
            t = 1.0 + abs(alpha) * rn
            scal = np.copysign(1.0 / (abs(alpha) + ab), alpha)
        an = A[:, k + 1] - (vi * wk1 + wi * vk1)       # a'_{i,k+1}: next pivot column after reflector k-1
        pn = t * (an + scal * ((sx - vi * wx) - wi * vx))
        p_next, r_next = np.zeros(n), np.zeros(n)
        p_next[k + 1:], r_next[k + 1:] = pn[k + 1:], an[k + 1:]          # rows i > k send
        # after the send: own rows catch up with reflector k-1; reflector k is normalised and recorded
        if k > 0:
            A[k:] -= np.outer(vi[k:], w) + np.outer(wi[k:], vprev)
        v = x * scal
        v[k + 1] = 1.0
        e[k], tau[k], V[k] = beta, t, v
        vprev, tprev, p_recv, r_recv = v, t, p_next, r_next
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
