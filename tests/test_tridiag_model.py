"""CPU model of the single-exchange Householder tridiagonalisation implemented by tridiag_cluster_kernel
(gpcsd_b200/csrc/gpcsd_eig.cu).  Per column every row owner sends the pair (p_i of the column just built, a_{i,k+1} = its
element of the NEXT pivot column); every participant then finishes the previous column (w = p - tau/2 (p.v) v), rebuilds
the current pivot row from the received elements, and builds the reflector with the kernel's formulas
(1 + |alpha|/|beta|, sign(alpha)/(|alpha|+|beta|)).  The symv runs on the UNNORMALISED pivot column and on rows that have
not yet been updated by the previous reflector (A' x = A x - v (w.x) - w (v.x)); what is sent is the RAW pair
((A' x)_i, a'_{i,k+1}) and the reflector is kept unnormalised, so that its scalars (rsqrt, reciprocal) are only needed one
exchange later, where their dependent chain runs next to the p.v reduction; the own rows catch up after the send.
The test pins that algebra against LAPACK."""
import numpy as np
import pytest
import scipy.linalg


def _reflector_scalars(alpha, xnorm2):
    """tau, beta, 1 / (alpha - beta) with the kernel's formulas"""
    if not xnorm2 > 0.0:
        return 0.0, alpha, 0.0
    s2 = alpha * alpha + xnorm2
    rn = 1.0 / np.sqrt(s2)
    ab = s2 * rn
    return 1.0 + abs(alpha) * rn, -np.copysign(ab, alpha), np.copysign(1.0 / (abs(alpha) + ab), alpha)


def tridiag_single_exchange(M):
    n = M.shape[0]
    A = M.copy()                      # rows as the owning warps hold them (updated lazily: one reflector behind at the symv)
    d, e, tau = np.zeros(n), np.zeros(n), np.zeros(n)
    V = np.zeros((n, n))
    xp, alpha_p, xn2_p = np.zeros(n), 0.0, 0.0        # UNNORMALISED reflector k-1 (zero for j <= k), its alpha and |x|^2
    s_recv, r_recv = np.zeros(n), A[:, 0].copy()      # exchange 0: column 0 of every row, raw sum = 0
    for k in range(n - 1):
        # ---- the two sums of the received raw pairs against the unnormalised reflector (no scalar of column k-1 needed) ...
        S1, S2 = float(r_recv @ xp), float(s_recv @ xp)
        # ---- ... while the rsqrt / reciprocal chain of column k-1 runs; column k-1 is recorded now
        t, beta, scal = _reflector_scalars(alpha_p, xn2_p)
        if k > 0:
            e[k - 1], tau[k - 1] = beta, t
            V[k - 1] = xp * scal
            V[k - 1, k] = 1.0
        # p_j = t (a_j + scal s_j), v_k = 1, v_j = scal x_j:  p.v = t (a_k + scal (s_k + S1 + scal S2))
        pv = t * (r_recv[k] + scal * (S1 + s_recv[k] + scal * S2))
        c = 0.5 * t * pv
        ts, cs = t * scal, c * scal
        w = np.zeros(n)
        w[k:] = t * r_recv[k:] + ts * s_recv[k:] - cs * xp[k:]      # p_j - c v_j  (exact at j > k; j = k handled by wk)
        pk = t * (r_recv[k] + scal * s_recv[k])
        wk = pk - c
        vi = scal * xp                                             # per own row i > k: v_i, w_i
        wi = t * (r_recv + scal * s_recv) - c * vi
        if k == n - 2:                 # last pass: rows n-2, n-1 by reflector n-3; here column k is NOT dead (v_k = 1)
            v = scal * xp
            v[k], w[k], wi[k] = 1.0, wk, wk
            A[k:] -= np.outer(v[k:], w) + np.outer(wi[k:], v)
            break
        vk1 = scal * xp[k + 1]
        wk1 = t * (r_recv[k + 1] + scal * s_recv[k + 1]) - c * vk1
        d[k] = r_recv[k] - 2.0 * wk
        alpha = r_recv[k + 1] - wk1 - wk * vk1
        x = np.zeros(n)
        x[k + 2:] = (r_recv[k + 2:] - w[k + 2:]) - (wk * scal) * xp[k + 2:]
        xnorm2, wx, xx = float(x @ x), float(w @ x), float(xp @ x)
        sx = A @ x                    # rows as held: BEFORE the update by reflector k-1
        an = A[:, k + 1] - (vi * wk1 + wi * vk1)       # a'_{i,k+1}: next pivot column after reflector k-1
        sp = (sx - vi * wx) - wi * (scal * xx)         # (A' x)_i
        s_next, r_next = np.zeros(n), np.zeros(n)
        s_next[k + 1:], r_next[k + 1:] = sp[k + 1:], an[k + 1:]          # rows i > k send
        # after the send: the own rows catch up with reflector k-1 (columns j > k; column k is dead)
        if k > 0:
            A[k + 1:, k + 1:] -= np.outer(vi[k + 1:], w[k + 1:]) + np.outer(wi[k + 1:] * scal, xp[k + 1:])
        xp, alpha_p, xn2_p, s_recv, r_recv = x, alpha, xnorm2, s_next, r_next
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
