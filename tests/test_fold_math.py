"""CPU checks of the symmetry-fold mathematics of DESIGN.md section 3.1 (no GPU, no library): the identities the folded
projection / SYRK / back-projection paths of gpcsd_b200.engine rely on, stated with numpy on oracle-built matrices."""
import numpy as np
import pytest


def _fold_matrix(n):
    """Orthogonal F with (F^T x) = [sums of mirrored entries / sqrt2 (middle kept for odd n); differences / sqrt2]
    -- the convention of gpcsd_centro_fold / centro_assemble_kernel."""
    m, ms = n // 2, n - n // 2
    F = np.zeros((n, n))
    h = 1.0 / np.sqrt(2.0)
    for j in range(m):
        F[j, j] = F[n - 1 - j, j] = h
        F[j, ms + j], F[n - 1 - j, ms + j] = h, -h
    if n % 2:
        F[m, m] = 1.0
    return F


@pytest.mark.parametrize("nt", [40, 41])
def test_time_fold_identities(nt):
    from oracle import synth
    rng = np.random.default_rng(nt)
    x, t = synth.geometry_1d(8, nt)
    om = synth.model_1d(x, t)
    Kt = om.Kt()
    m, ms = nt // 2, nt - nt // 2
    F = _fold_matrix(nt)
    assert np.allclose(F.T @ F, np.eye(nt), atol=1e-15)
    # centrosymmetric Kt is block diagonal in the folded basis; the blocks are S = K11 + K12 J, A = K11 - K12 J
    B = F.T @ Kt @ F
    assert np.max(np.abs(B[:ms, ms:])) < 1e-14 * np.max(np.abs(Kt))
    J = np.eye(m)[::-1]
    assert np.allclose(B[ms:, ms:], Kt[:m, :m] - Kt[:m, nt - m:] @ J, atol=1e-13)
    # eigenvectors: Qt = F blockdiag(Us, Ua), so Qt^T z = [Us^T (F^T z)_s ; Ua^T (F^T z)_a]
    ws, Us = np.linalg.eigh(B[:ms, :ms])
    wa, Ua = np.linalg.eigh(B[ms:, ms:])
    Qt = F @ np.block([[Us, np.zeros((ms, m))], [np.zeros((m, ms)), Ua]])
    lam = np.concatenate([ws, wa])
    assert np.max(np.abs(Kt @ Qt - Qt * lam)) < 1e-12 * np.max(np.abs(Kt))
    z = rng.standard_normal((nt, 5))
    zf = F.T @ z
    assert np.allclose(Qt.T @ z, np.concatenate([Us.T @ zf[:ms], Ua.T @ zf[ms:]]), atol=1e-12)
    # gradient: <Qt M Qt^T, C> only sees the diagonal blocks of M for every centrosymmetric C (all dKt/dtheta, Kt*_k at t*=t)
    M = rng.standard_normal((nt, nt))
    M = M + M.T
    Mb = M.copy()
    Mb[:ms, ms:] = 0.0
    Mb[ms:, :ms] = 0.0
    dt = t - t.T
    for C in (np.exp(-0.5 * dt ** 2 / 9.0) * dt ** 2, np.exp(-np.abs(dt) / 3.0) * np.abs(dt), Kt):
        full, blocks = np.sum((Qt @ M @ Qt.T) * C), np.sum((Qt @ Mb @ Qt.T) * C)
        assert abs(full - blocks) <= 1e-11 * max(abs(full), np.max(np.abs(C)) * np.max(np.abs(M)) * nt)
    # predict's back-projection: C Qt V = F blockdiag(Cs Us, Ca Ua) V
    V = rng.standard_normal((nt, 4))
    Cf = F.T @ Kt @ F
    rhs = F @ np.concatenate([Cf[:ms, :ms] @ Us @ V[:ms], Cf[ms:, ms:] @ Ua @ V[ms:]])
    assert np.allclose(Kt @ Qt @ V, rhs, atol=1e-11 * np.max(np.abs(Kt)) * nt)


def test_channel_fold_identities():
    """Reflection-symmetric probe geometry: Ks and every dKs/dtheta commute with the site pairing, hence are block
    diagonal in the channel-folded basis of gpcsd_pairsym_fold."""
    from oracle import gpcsd_oracle as O, synth
    X, t = synth.geometry_neuropixels(32, 8, 0.4)
    ymax = X[:, 1].max()
    om = synth.model_2d(X, t, ngl1=6, ngl2=16, a1=-16.0, b1=64.0, a2=-100.0, b2=ymax + 100.0, eps=1.0, sig2n=0.5)
    Ks = om.Ks(jitter=True)
    nx = Ks.shape[0]
    c = np.array([-16.0 + 64.0, -100.0 + ymax + 100.0])
    Xr = c - X                                      # point reflection about the centre of the integration box
    pi = np.array([int(np.argmin(np.sum((X - p) ** 2, axis=1))) for p in Xr])
    assert np.max(np.abs(X[pi] - Xr)) < 1e-9 and np.all(pi[pi] == np.arange(nx)) and np.all(pi != np.arange(nx))
    assert np.max(np.abs(Ks[np.ix_(pi, pi)] - Ks)) < 1e-12 * np.max(np.abs(Ks))
    ra = np.array([i for i in range(nx) if i < pi[i]])
    rb = pi[ra]
    mh = nx // 2
    F = np.zeros((nx, nx))
    h = 1.0 / np.sqrt(2.0)
    for cidx, (a, b) in enumerate(zip(ra, rb)):
        F[a, cidx] = F[b, cidx] = h
        F[a, mh + cidx], F[b, mh + cidx] = h, -h
    B = F.T @ Ks @ F
    assert np.max(np.abs(B[:mh, mh:])) < 1e-12 * np.max(np.abs(Ks))
    # a perturbed-hyperparameter Ks (stands for any dKs/dtheta by finite differences) keeps the block structure
    om2 = synth.perturbed(om, 3, scale=0.2)
    B2 = F.T @ om2.Ks(jitter=False) @ F
    assert np.max(np.abs(B2[:mh, mh:])) < 1e-12 * np.max(np.abs(B2))


def _dct2(n):
    i, k = np.arange(n)[:, None], np.arange(n)[None, :]
    return np.sqrt(np.where(k == 0, 1.0, 2.0) / n) * np.cos(np.pi * (2 * i + 1) * k / (2 * n))


def _dct4(n):
    i, k = np.arange(n)[:, None], np.arange(n)[None, :]
    return np.sqrt(2.0 / n) * np.cos(np.pi * (2 * i + 1) * (2 * k + 1) / (4 * n))


@pytest.mark.parametrize("m", [12, 25])
def test_cosine_prerotation_of_the_centrosymmetric_halves(m):
    """jacobi_small_kernel (gpcsd_eig.cu) pre-rotates by DCT-II or DCT-IV, whichever leaves the larger squared diagonal.
    Both are orthonormal; the even / odd members of the DCT-II basis of order 2m restricted to the first half ARE the DCT-II /
    DCT-IV bases of order m (up to the fold's sqrt 2), so the symmetric half of a stationary kernel's centrosymmetric split is
    closest to diagonal in the first basis and the skew half in the second -- the selection rule picks them that way."""
    n = 2 * m
    C2, C4, Cn = _dct2(m), _dct4(m), _dct2(n)
    assert np.max(np.abs(C2.T @ C2 - np.eye(m))) < 1e-13 and np.max(np.abs(C4.T @ C4 - np.eye(m))) < 1e-13
    assert np.max(np.abs(np.sqrt(2.0) * Cn[:m, 0::2] - C2)) < 1e-13
    assert np.max(np.abs(np.sqrt(2.0) * Cn[:m, 1::2] - C4)) < 1e-13
    t = np.arange(n, dtype=float)
    dd = t[:, None] - t[None, :]
    K = 0.5 * np.exp(-0.5 * dd ** 2 / 40.0) + 0.7 * np.exp(-np.abs(dd) / 5.0)
    J = np.eye(m)[::-1]
    S, A = K[:m, :m] + K[:m, m:] @ J, K[:m, :m] - K[:m, m:] @ J          # the two halves of F^T K F
    score = lambda M, C: float(np.sum(np.diag(C.T @ M @ C) ** 2))
    assert score(S, C2) > score(S, C4)
    assert score(A, C4) > score(A, C2)
    off = lambda M, C: np.linalg.norm(C.T @ M @ C - np.diag(np.diag(C.T @ M @ C))) / np.linalg.norm(M)
    assert off(S, C2) < 0.25 * off(S, np.eye(m)) and off(A, C4) < 0.25 * off(A, np.eye(m))


def test_transposed_eight_value_warp_reduction_model():
    """warp_reduce8_transposed (gpcsd_eig.cu): every stage halves the values a lane carries; lane L ends with the warp total of
    value L >> 2, and the four lanes of a group agree bit for bit (partners add the same two numbers)."""
    rng = np.random.default_rng(0)
    v = rng.standard_normal((32, 8))                       # v[lane][i]
    lanes = np.arange(32)
    shfl = lambda x, o: x[lanes ^ o]
    b4, b3, b2 = (lanes & 16) != 0, (lanes & 8) != 0, (lanes & 4) != 0
    u = np.stack([np.where(b4, v[:, i + 4], v[:, i]) + shfl(np.where(b4, v[:, i], v[:, i + 4]), 16) for i in range(4)], axis=1)
    w = np.stack([np.where(b3, u[:, i + 2], u[:, i]) + shfl(np.where(b3, u[:, i], u[:, i + 2]), 8) for i in range(2)], axis=1)
    r = np.where(b2, w[:, 1], w[:, 0]) + shfl(np.where(b2, w[:, 0], w[:, 1]), 4)
    r = r + shfl(r, 2)
    r = r + shfl(r, 1)
    tot = v.sum(axis=0)
    for L in range(32):
        assert abs(r[L] - tot[L >> 2]) < 1e-13
        assert r[L] == r[L & ~3]
