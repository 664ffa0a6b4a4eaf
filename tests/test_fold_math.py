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
