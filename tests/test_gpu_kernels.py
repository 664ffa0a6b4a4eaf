"""Kernel-level parity of the C-ABI entry points against numpy / the oracle pieces (-m gpu)."""
import numpy as np
import pytest
import torch

from helpers import relerr

pytestmark = pytest.mark.gpu
F64 = torch.float64


def _ld(n):
    return (n + 1) // 2 * 2


def _dev(a, ld=None):
    """row-major matrix -> device buffer with even leading dimension; returns (tensor, ld)."""
    a = np.atleast_2d(np.asarray(a, dtype=np.float64))
    ld = _ld(a.shape[1]) if ld is None else ld
    buf = torch.zeros((a.shape[0], ld), dtype=F64, device="cuda")
    buf[:, : a.shape[1]] = torch.from_numpy(a).cuda()
    return buf, ld


def _stream():
    return torch.cuda.current_stream().cuda_stream


@pytest.mark.parametrize("M,N,K,batch", [(128, 128, 16, 1), (500, 2000, 500, 3), (24, 1003, 24, 1), (7, 50, 5, 2),
                                          (33, 65, 17, 2), (384, 777, 384, 1), (130, 258, 131, 1), (1, 9, 300, 1),
                                          (24, 100008, 24, 1), (32, 4101, 32, 2), (9, 17, 31, 3), (24, 24, 24, 1), (30, 8, 3, 1),
                                          # 64-row TMA tiles (M pads badly to 128), latency mode (few 128 x 64 tiles)
                                          (192, 3000, 192, 2), (150, 2100, 70, 1), (250, 250, 250, 2), (192, 192, 192, 2)])
@pytest.mark.parametrize("transB", [0, 1])
def test_dgemm(cuda_lib, M, N, K, batch, transB):
    from gpcsd_b200 import _lib as L
    rng = np.random.default_rng(M * 7 + N + K + transB)
    lda, ldb, ldc = _ld(K), _ld(K if transB else N), _ld(N)
    A = torch.zeros((batch, M, lda), dtype=F64, device="cuda")
    B = torch.zeros((batch, N if transB else K, ldb), dtype=F64, device="cuda")
    C = torch.full((batch, M, ldc), 7.0, dtype=F64, device="cuda")
    Ah = rng.standard_normal((batch, M, K))
    Bh = rng.standard_normal((batch, N, K) if transB else (batch, K, N))
    A[:, :, :K] = torch.from_numpy(Ah).cuda()
    B[:, :, : Bh.shape[2]] = torch.from_numpy(Bh).cuda()
    L.call("gpcsd_dgemm", transB, M, N, K, A.data_ptr(), lda, M * lda, B.data_ptr(), ldb, B.shape[1] * ldb,
           C.data_ptr(), ldc, M * ldc, batch, _stream())
    torch.cuda.synchronize()
    ref = np.einsum("bmk,bnk->bmn", Ah, Bh) if transB else np.einsum("bmk,bkn->bmn", Ah, Bh)
    got = C[:, :, :N].cpu().numpy()
    assert relerr(got, ref) < 1e-13
    if ldc > N:  # padding column untouched
        assert torch.all(C[:, :, N:] == 7.0)


def test_dgemm_broadcast_A_and_strided_view(cuda_lib):
    """strideA == 0 (shared small matrix) and B viewed as [nx][nt*ldn] exactly as the engine uses it."""
    from gpcsd_b200 import _lib as L
    rng = np.random.default_rng(3)
    nx, nt, ldn = 5, 37, 24
    Q = rng.standard_normal((nt, nt))
    Y = rng.standard_normal((nx, nt, ldn))
    Qd, ldq = _dev(Q)
    Yd = torch.from_numpy(Y).cuda()
    Cd = torch.zeros_like(Yd)
    L.call("gpcsd_dgemm", 0, nt, ldn, nt, Qd.data_ptr(), ldq, 0, Yd.data_ptr(), ldn, nt * ldn, Cd.data_ptr(), ldn,
           nt * ldn, nx, _stream())
    assert relerr(Cd.cpu().numpy(), np.einsum("ab,ibr->iar", Q, Y)) < 1e-13


@pytest.mark.parametrize("nx,nt,N", [(24, 50, 50), (6, 130, 257), (3, 20, 1), (40, 33, 17)])
def test_project_quad(cuda_lib, nx, nt, N):
    from gpcsd_b200 import _lib as L
    rng = np.random.default_rng(nx + nt + N)
    ldn = (N + 7) // 8 * 8
    QtT = rng.standard_normal((nt, nt))
    Z = np.zeros((nx, nt, ldn))
    Z[:, :, :N] = rng.standard_normal((nx, nt, N))
    rD = rng.uniform(0.5, 2.0, (nx, nt))
    Qd, ldq = _dev(QtT)
    rDd, ldrd = _dev(rD)
    Zd = torch.from_numpy(Z).cuda()
    Bd = torch.zeros_like(Zd)
    part = torch.zeros(L.query("gpcsd_project_quad_ws_doubles", nx, nt, N), dtype=F64, device="cuda")
    out2 = torch.zeros(2, dtype=F64, device="cuda")
    L.call("gpcsd_project_quad", nx, nt, N, Qd.data_ptr(), ldq, Zd.data_ptr(), ldn, rDd.data_ptr(), ldrd,
           Bd.data_ptr(), part.data_ptr(), out2.data_ptr(), _stream())
    A = np.einsum("ab,ibr->iar", QtT, Z[:, :, :N])
    Bm = A * rD[:, :, None]
    assert relerr(Bd[:, :, :N].cpu().numpy(), Bm) < 1e-13
    o = out2.cpu().numpy()
    assert abs(o[0] - np.sum(A * Bm)) / np.sum(A * Bm) < 1e-13
    assert abs(o[1] - np.sum(Bm * Bm)) / np.sum(Bm * Bm) < 1e-13


@pytest.mark.parametrize("nx,nt,N", [(24, 50, 50), (5, 140, 300), (130, 9, 33), (3, 3, 1), (192, 7, 40), (6, 250, 64), (64, 3, 24)])
@pytest.mark.parametrize("weighted", [True, False])
def test_wsyrk_both_orientations(cuda_lib, nx, nt, N, weighted):
    from gpcsd_b200 import _lib as L
    rng = np.random.default_rng(nx * 3 + nt + N)
    ldn = (N + 7) // 8 * 8
    Bm = np.zeros((nx, nt, ldn))
    Bm[:, :, :N] = rng.standard_normal((nx, nt, N))
    Bm[:, :, N:] = 99.0  # garbage in the padding must be ignored
    ls, lt = rng.standard_normal(nx), rng.standard_normal(nt)
    Bd = torch.from_numpy(Bm).cuda()
    lsd, ltd = torch.from_numpy(ls).cuda(), torch.from_numpy(lt).cuda()
    B = Bm[:, :, :N]
    # Mt: M = nt, segments = nx
    ldt = _ld(nt)
    Mt = torch.zeros((nt, ldt), dtype=F64, device="cuda")
    ws = torch.zeros(L.query("gpcsd_wsyrk_ws_doubles", nt, nx, N), dtype=F64, device="cuda")
    L.call("gpcsd_wsyrk", nt, nx, N, Bd.data_ptr(), ldn, nt * ldn, lsd.data_ptr() if weighted else None, Mt.data_ptr(),
           ldt, ws.data_ptr(), _stream())
    ref = np.einsum("ajr,a,akr->jk", B, ls if weighted else np.ones(nx), B)
    assert relerr(Mt[:, :nt].cpu().numpy(), ref) < 1e-12
    # Ms: M = nx, segments = nt
    ldx = _ld(nx)
    Ms = torch.zeros((nx, ldx), dtype=F64, device="cuda")
    ws = torch.zeros(L.query("gpcsd_wsyrk_ws_doubles", nx, nt, N), dtype=F64, device="cuda")
    L.call("gpcsd_wsyrk", nx, nt, N, Bd.data_ptr(), nt * ldn, ldn, ltd.data_ptr() if weighted else None, Ms.data_ptr(),
           ldx, ws.data_ptr(), _stream())
    ref = np.einsum("ajr,j,bjr->ab", B, lt if weighted else np.ones(nt), B)
    assert relerr(Ms[:, :nx].cpu().numpy(), ref) < 1e-12


@pytest.mark.parametrize("nx,nt,N", [(24, 50, 50), (24, 500, 333), (5, 140, 300), (130, 9, 33), (3, 3, 1)])
def test_wsyrk_pair_one_pass(cuda_lib, nx, nt, N):
    """Ms = sum_j lt_j B_j B_j^T and Ns = sum_j B_j B_j^T from one pass over Bm (nx <= 32) or two (nx > 32)."""
    from gpcsd_b200 import _lib as L
    rng = np.random.default_rng(nx + 7 * nt + N)
    ldn = (N + 7) // 8 * 8
    Bm = np.zeros((nx, nt, ldn))
    Bm[:, :, :N] = rng.standard_normal((nx, nt, N))
    Bm[:, :, N:] = 99.0
    lt = rng.standard_normal(nt)
    Bd, ltd = torch.from_numpy(Bm).cuda(), torch.from_numpy(lt).cuda()
    B = Bm[:, :, :N]
    ldx = _ld(nx)
    Ms = torch.zeros((nx, ldx), dtype=F64, device="cuda")
    Ns = torch.zeros((nx, ldx), dtype=F64, device="cuda")
    ws = torch.zeros(2 * L.query("gpcsd_wsyrk_ws_doubles", nx, nt, N), dtype=F64, device="cuda")
    L.call("gpcsd_wsyrk_pair", nx, nt, N, Bd.data_ptr(), nt * ldn, ldn, ltd.data_ptr(), Ms.data_ptr(), Ns.data_ptr(), ldx,
           ws.data_ptr(), _stream())
    assert relerr(Ms[:, :nx].cpu().numpy(), np.einsum("ajr,j,bjr->ab", B, lt, B)) < 1e-12
    assert relerr(Ns[:, :nx].cpu().numpy(), np.einsum("ajr,bjr->ab", B, B)) < 1e-12


@pytest.mark.parametrize("R,nx,nt,N", [(3, 24, 50, 50), (2, 24, 250, 96), (5, 5, 20, 9), (2, 40, 130, 33)])
def test_batched_projection_and_syrk(cuda_lib, R, nx, nt, N):
    """Restart-batched forms (one launch for R restarts) of the projection with the /D + quadratic-form epilogue and of the
    segment-weighted SYRKs, against numpy per restart."""
    from gpcsd_b200 import _lib as L
    rng = np.random.default_rng(R * 100 + nx + nt + N)
    ldn, ldt, ldx = (N + 7) // 8 * 8, _ld(nt), _ld(nx)
    Z = np.zeros((R, nx, nt, ldn)); Z[..., :N] = rng.standard_normal((R, nx, nt, N)); Z[..., N:] = 55.0
    Q = rng.standard_normal((R, nt, nt))
    rD = rng.uniform(0.5, 2.0, (R, nx, nt))
    ls, lt = rng.standard_normal((R, nx)), rng.standard_normal((R, nt))
    Zd = torch.from_numpy(Z).cuda()
    QT = torch.zeros((R, nt, ldt), dtype=F64, device="cuda"); QT[:, :, :nt] = torch.from_numpy(Q).cuda()
    rDd = torch.zeros((R, nx, ldt), dtype=F64, device="cuda"); rDd[:, :, :nt] = torch.from_numpy(rD).cuda()
    Bd = torch.zeros_like(Zd)
    part = torch.zeros(L.query("gpcsd_project_quad_batched_ws_doubles", R, nx, nt, N), dtype=F64, device="cuda")
    out = torch.zeros((R, 32), dtype=F64, device="cuda")
    L.call("gpcsd_project_quad_batched", R, nx, nt, N, QT.data_ptr(), ldt, nt * ldt, Zd.data_ptr(), ldn, nt * ldn, rDd.data_ptr(), ldt,
           Bd.data_ptr(), part.data_ptr(), out.data_ptr(), 32, _stream())
    A = np.einsum("rab,ribn->rian", Q, Z[..., :N])           # A_i = QT Z_i per restart (rows of QT = eigenvectors)
    Bm = A * rD[..., None]
    assert relerr(Bd[..., :N].cpu().numpy(), Bm) < 1e-13
    o = out.cpu().numpy()
    for r in range(R):
        assert abs(o[r, 0] - np.sum(A[r] * Bm[r])) < 1e-12 * np.sum(np.abs(A[r] * Bm[r]))
        assert abs(o[r, 1] - np.sum(Bm[r] ** 2)) < 1e-12 * np.sum(Bm[r] ** 2)
    # SYRKs on Bd (padding columns hold zeros there)
    lsd, ltd = torch.from_numpy(ls).cuda(), torch.from_numpy(lt).cuda()
    Mt = torch.zeros((R, nt, ldt), dtype=F64, device="cuda")
    ws = torch.zeros(L.query("gpcsd_wsyrk_batched_ws_doubles", R, nt, nx, N, 0), dtype=F64, device="cuda")
    L.call("gpcsd_wsyrk_batched", R, nt, nx, N, Bd.data_ptr(), ldn, nt * ldn, nx * nt * ldn, lsd.data_ptr(), nx, Mt.data_ptr(), None, ldt,
           nt * ldt, ws.data_ptr(), _stream())
    assert relerr(Mt[:, :, :nt].cpu().numpy(), np.einsum("rajn,ra,rakn->rjk", Bm, ls, Bm)) < 1e-12
    Ms = torch.zeros((R, nx, ldx), dtype=F64, device="cuda")
    Ns = torch.zeros((R, nx, ldx), dtype=F64, device="cuda")
    ws = torch.zeros(L.query("gpcsd_wsyrk_batched_ws_doubles", R, nx, nt, N, 1), dtype=F64, device="cuda")
    L.call("gpcsd_wsyrk_batched", R, nx, nt, N, Bd.data_ptr(), nt * ldn, ldn, nx * nt * ldn, ltd.data_ptr(), nt, Ms.data_ptr(), Ns.data_ptr(),
           ldx, nx * ldx, ws.data_ptr(), _stream())
    assert relerr(Ms[:, :, :nx].cpu().numpy(), np.einsum("rajn,rj,rbjn->rab", Bm, lt, Bm)) < 1e-12
    assert relerr(Ns[:, :, :nx].cpu().numpy(), np.einsum("rajn,rbjn->rab", Bm, Bm)) < 1e-12


@pytest.mark.parametrize("n", [5, 24, 100, 251])
def test_eigh(cuda_lib, n):
    from gpcsd_b200 import _lib as L
    rng = np.random.default_rng(n)
    K = rng.standard_normal((n, n))
    K = K @ K.T
    Kd, ld = _dev(K)
    QT = torch.zeros((n, ld), dtype=F64, device="cuda")
    W = torch.zeros(n, dtype=F64, device="cuda")
    nws = L.query("gpcsd_eigh_ws_doubles", n, ld)
    ws = torch.zeros(max(nws, 1), dtype=F64, device="cuda")
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    L.call("gpcsd_eigh", n, Kd.data_ptr(), ld, QT.data_ptr(), ld, W.data_ptr(), ws.data_ptr(), nws, info.data_ptr(), _stream())
    assert int(info.item()) == 0
    w = W.cpu().numpy()
    Q = QT[:, :n].cpu().numpy().T
    assert np.all(np.diff(w) >= 0)
    assert relerr(w, np.linalg.eigvalsh(K)) < 1e-12
    assert relerr((Q * w) @ Q.T, K) < 1e-12
    assert relerr(Q.T @ Q, np.eye(n)) < 1e-12


@pytest.mark.parametrize("vec", [False, True])
def test_eig_D(cuda_lib, vec):
    from gpcsd_b200 import _lib as L
    rng = np.random.default_rng(1)
    nx, nt = 24, 77
    ls, lt = np.sort(rng.uniform(0, 5, nx)), np.sort(rng.uniform(0, 2, nt))
    s = rng.uniform(0.1, 0.3, nx) if vec else np.array([0.2])
    ldrd = _ld(nt)
    rD = torch.zeros((nx, ldrd), dtype=F64, device="cuda")
    sums, rowA, rowC, rowL = (torch.zeros(k, dtype=F64, device="cuda") for k in (2, nx, nx, nx))
    colB = torch.zeros(nt, dtype=F64, device="cuda")
    d = lambda a: torch.from_numpy(a).cuda()
    lsd, ltd, sd = d(ls), d(lt), d(s)
    L.call("gpcsd_eig_D", nx, nt, lsd.data_ptr(), ltd.data_ptr(), sd.data_ptr(), len(s), rD.data_ptr(), ldrd,
           sums.data_ptr(), rowA.data_ptr(), rowC.data_ptr(), rowL.data_ptr(), colB.data_ptr(), _stream())
    D = ls[:, None] * lt[None, :] + (s[:, None] if vec else s[0])
    assert relerr(rD[:, :nt].cpu().numpy(), 1 / D) < 1e-14
    assert relerr(sums.cpu().numpy(), [np.sum(np.log(D)), np.sum(1 / D)]) < 1e-13
    assert relerr(rowA.cpu().numpy(), (lt[None, :] / D).sum(1)) < 1e-13
    assert relerr(rowC.cpu().numpy(), (1 / D).sum(1)) < 1e-13
    assert relerr(rowL.cpu().numpy(), np.log(D).sum(1)) < 1e-13
    assert relerr(colB.cpu().numpy(), (ls[:, None] / D).sum(0)) < 1e-13


def test_covariance_builders_match_oracle(cuda_lib):
    from gpcsd_b200 import _lib as L
    from oracle import gpcsd_oracle as O
    rng = np.random.default_rng(2)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()
    # 1-D forward weights + SE matrices
    x = np.linspace(0, 2300, 24)
    gx, gw = O.gauss_legendre(-200.0, 2600.0, 100)
    A = torch.zeros((24, 100), dtype=F64, device="cuda")
    dA = torch.zeros_like(A)
    xd, gxd, gwd = d(x), d(gx), d(gw)
    L.call("gpcsd_fwd_weights_1d", 24, xd.data_ptr(), 100, gxd.data_ptr(), gwd.data_ptr(), 137.0, A.data_ptr(),
           dA.data_ptr(), 100, _stream())
    ref = gw[None, :] * O.b_fwd_1d(gx[None, :] - x[:, None], 137.0)
    assert relerr(A.cpu().numpy(), ref) < 1e-14
    h = 1e-3
    fd = gw[None, :] * (O.b_fwd_1d(gx[None, :] - x[:, None], 137.0 + h) - O.b_fwd_1d(gx[None, :] - x[:, None], 137.0 - h)) / (2 * h)
    assert relerr(dA.cpu().numpy(), fd) < 1e-7
    Kg = torch.zeros((100, 100), dtype=F64, device="cuda")
    L.call("gpcsd_se_matrix", 100, gxd.data_ptr(), 100, gxd.data_ptr(), 200.0, 1.0, 0, Kg.data_ptr(), 100, _stream())
    assert relerr(Kg.cpu().numpy(), np.exp(-0.5 * np.square((gx[:, None] - gx[None, :]) / 200.0))) < 1e-14
    L.call("gpcsd_se_matrix", 100, gxd.data_ptr(), 100, gxd.data_ptr(), 200.0, 1.0, 1, Kg.data_ptr(), 100, _stream())
    dd = gx[:, None] - gx[None, :]
    assert relerr(Kg.cpu().numpy(), np.exp(-0.5 * np.square(dd / 200.0)) * dd ** 2 / 200.0 ** 3) < 1e-13
    # 2-D forward weights
    pts = rng.uniform(0, 50, (7, 2))
    g1, w1 = O.gauss_legendre(0.0, 48.0, 6)
    g2, w2 = O.gauss_legendre(0.0, 220.0, 10)
    sp = O.Spatial2D(pts, 0.0, 48.0, 0.0, 220.0, 6, 10)
    A2 = torch.zeros((7, 60), dtype=F64, device="cuda")
    dA2 = torch.zeros_like(A2)
    pd_, g1d, w1d, g2d, w2d = d(pts), d(g1), d(w1), d(g2), d(w2)
    L.call("gpcsd_fwd_weights_2d", 7, pd_.data_ptr(), 6, 10, g1d.data_ptr(), w1d.data_ptr(), g2d.data_ptr(),
           w2d.data_ptr(), 60.0, 20.0, A2.data_ptr(), dA2.data_ptr(), 60, _stream())
    ref2 = sp.w_prod[None, :] * O.b_fwd_2d(sp.delta_w(pts), 60.0, 20.0)
    assert relerr(A2.cpu().numpy(), ref2) < 1e-14
    fd2 = sp.w_prod[None, :] * (O.b_fwd_2d(sp.delta_w(pts), 60.0 + h, 20.0) - O.b_fwd_2d(sp.delta_w(pts), 60.0 - h, 20.0)) / (2 * h)
    assert relerr(dA2.cpu().numpy(), fd2) < 1e-7
    # grid-to-points kernel (transposed storage)
    z = rng.uniform(0, 50, (5, 2))
    out = torch.zeros((5, 60), dtype=F64, device="cuda")
    zd = d(z)
    L.call("gpcsd_se_grid_to_pts", 6, 10, g1d.data_ptr(), g2d.data_ptr(), 5, zd.data_ptr(), 30.0, 70.0, out.data_ptr(), 60, _stream())
    refz = (np.exp(-0.5 * np.square((sp.grid1[:, None] - z[:, 0][None, :]) / 30.0))
            * np.exp(-0.5 * np.square((sp.grid2[:, None] - z[:, 1][None, :]) / 70.0)))
    assert relerr(out.cpu().numpy(), refz.T) < 1e-14
    # temporal covariance, rectangular
    t, tp = np.linspace(0, 49, 50), np.linspace(0.5, 30, 31)
    Kt = torch.zeros((50, 32), dtype=F64, device="cuda")
    td, tpd = d(t), d(tp)
    L.call("gpcsd_kt_build", 50, td.data_ptr(), 31, tpd.data_ptr(), 2, L.c_int_array([0, 1]), L.c_double_array([20.0, 5.0]),
           L.c_double_array([0.5, 0.7]), Kt.data_ptr(), 32, _stream())
    ref = O.compute_Kt(0, 20.0, 0.5, t, tp) + O.compute_Kt(1, 5.0, 0.7, t, tp)
    assert relerr(Kt[:, :31].cpu().numpy(), ref) < 1e-14


def test_kt_grad_and_small_helpers(cuda_lib):
    from gpcsd_b200 import _lib as L
    from oracle import gpcsd_oracle as O
    rng = np.random.default_rng(4)
    nt = 61
    t = np.sort(rng.uniform(0, 80, nt))
    G = rng.standard_normal((nt, nt))
    Gd, ldg = _dev(G)
    td = torch.from_numpy(t).cuda()
    ws = torch.zeros(L.query("gpcsd_kt_grad_ws_doubles", nt, 2), dtype=F64, device="cuda")
    out = torch.zeros(4, dtype=F64, device="cuda")
    L.call("gpcsd_kt_grad", nt, td.data_ptr(), 2, L.c_int_array([0, 1]), L.c_double_array([20.0, 5.0]),
           L.c_double_array([0.5, 0.7]), Gd.data_ptr(), ldg, ws.data_ptr(), out.data_ptr(), _stream())
    dist = t[:, None] - t[None, :]
    Kse, Km = O.compute_Kt(0, 20.0, 0.5, t), O.compute_Kt(1, 5.0, 0.7, t)
    ref = [np.sum(G * Kse * dist ** 2 / 20.0 ** 3), np.sum(G * Kse) / 0.5, np.sum(G * Km * np.abs(dist) / 25.0), np.sum(G * Km) / 0.7]
    assert relerr(out.cpu().numpy(), ref) < 1e-12
    # transpose / add_diag / dot / sum_arrays
    X = rng.standard_normal((13, 40))
    Xd, ldx = _dev(X)
    XT = torch.zeros((40, 14), dtype=F64, device="cuda")
    L.call("gpcsd_transpose", 13, 40, Xd.data_ptr(), ldx, XT.data_ptr(), 14, _stream())
    assert np.array_equal(XT[:, :13].cpu().numpy(), X.T)
    Yh = rng.standard_normal((13, 40))
    Yd, ldy = _dev(Yh)
    wsd = torch.zeros(L.query("gpcsd_dot_ws_doubles", 13 * 40), dtype=F64, device="cuda")
    o1 = torch.zeros(1, dtype=F64, device="cuda")
    L.call("gpcsd_dot", 13, 40, Xd.data_ptr(), ldx, Yd.data_ptr(), ldy, wsd.data_ptr(), o1.data_ptr(), _stream())
    assert abs(o1.item() - np.sum(X * Yh)) < 1e-12 * np.sum(np.abs(X * Yh))
    K = torch.zeros((6, 6), dtype=F64, device="cuda")
    L.call("gpcsd_add_diag", 6, K.data_ptr(), 6, 1e-8, _stream())
    assert np.array_equal(K.cpu().numpy(), 1e-8 * np.eye(6))
    import ctypes
    a, b = torch.from_numpy(rng.standard_normal(1000)).cuda(), torch.from_numpy(rng.standard_normal(1000)).cuda()
    o = torch.zeros(1000, dtype=F64, device="cuda")
    ptrs = (ctypes.c_void_p * 2)(a.data_ptr(), b.data_ptr())
    L.call("gpcsd_sum_arrays", 1000, 2, ptrs, o.data_ptr(), _stream())
    assert np.array_equal(o.cpu().numpy(), (a + b).cpu().numpy())


@pytest.mark.parametrize("n", [2, 7, 40, 41, 250, 501])
def test_centrosymmetric_split_is_exact(cuda_lib, n):
    """S/A blocks of a symmetric Toeplitz matrix -> eigenpairs of the halves -> assembled QT, W reproduce K."""
    from gpcsd_b200 import _lib as L
    t = np.arange(n, dtype=np.float64) * 0.4
    d = t[:, None] - t[None, :]
    K = 0.5 * np.exp(-0.5 * d ** 2 / 25.0) + 0.7 * np.exp(-np.abs(d) / 3.0)
    m, ms = n // 2, n // 2 + (n & 1)
    Kd, ld = _dev(K)
    lds, lda = _ld(ms), _ld(max(m, 1))
    S = torch.zeros((ms, lds), dtype=F64, device="cuda")
    A = torch.zeros((max(m, 1), lda), dtype=F64, device="cuda")
    L.call("gpcsd_centro_split", n, Kd.data_ptr(), ld, S.data_ptr(), lds, A.data_ptr(), lda, _stream())
    J = np.eye(m)[::-1]
    Sh = S[:, :ms].cpu().numpy()
    assert relerr(Sh[:m, :m], K[:m, :m] + K[:m, n - m:] @ J) < 1e-15
    assert relerr(A[:m, :m].cpu().numpy(), K[:m, :m] - K[:m, n - m:] @ J) < 1e-15

    def eig(Md, k, ldm):
        QT = torch.zeros((k, ldm), dtype=F64, device="cuda")
        W = torch.zeros(k, dtype=F64, device="cuda")
        nws = L.query("gpcsd_eigh_ws_doubles", k, ldm)
        ws = torch.zeros(max(nws, 1), dtype=F64, device="cuda")
        info = torch.zeros(1, dtype=torch.int32, device="cuda")
        L.call("gpcsd_eigh", k, Md.data_ptr(), ldm, QT.data_ptr(), ldm, W.data_ptr(), ws.data_ptr(), nws, info.data_ptr(), _stream())
        assert int(info.item()) == 0
        return QT, W

    UsT, Ws = eig(S, ms, lds)
    UaT, Wa = eig(A, m, lda)
    QT = torch.zeros((n, ld), dtype=F64, device="cuda")
    W = torch.zeros(n, dtype=F64, device="cuda")
    L.call("gpcsd_centro_assemble", n, UsT.data_ptr(), lds, Ws.data_ptr(), UaT.data_ptr(), lda, Wa.data_ptr(),
           QT.data_ptr(), ld, W.data_ptr(), _stream())
    Q = QT[:, :n].cpu().numpy().T
    w = W.cpu().numpy()
    assert relerr(Q.T @ Q, np.eye(n)) < 1e-13
    assert relerr((Q * w) @ Q.T, K) < 1e-13
    assert relerr(np.sort(w), np.linalg.eigvalsh(K)) < 1e-12


@pytest.mark.parametrize("n,batch", [(24, 2), (50, 5), (125, 2), (128, 3)])
def test_eigh_batched_small_orders(cuda_lib, n, batch):
    from gpcsd_b200 import _lib as L
    rng = np.random.default_rng(n + batch)
    ld = _ld(n)
    Ks = []
    stack = torch.zeros((batch, n, ld), dtype=F64, device="cuda")
    for b in range(batch):
        K = rng.standard_normal((n, n))
        K = K @ K.T + b * np.eye(n)
        Ks.append(K)
        stack[b, :, :n] = torch.from_numpy(K).cuda()
    W = torch.zeros((batch, n), dtype=F64, device="cuda")
    nbytes = L.query("gpcsd_eigh_batched_ws_bytes", n, ld, batch)
    ws = torch.zeros(max((nbytes + 7) // 8, 1), dtype=F64, device="cuda")
    info = torch.ones(batch, dtype=torch.int32, device="cuda")
    L.call("gpcsd_eigh_batched", n, batch, stack.data_ptr(), ld, W.data_ptr(), ws.data_ptr(), nbytes, info.data_ptr(), _stream())
    assert torch.all(info == 0)
    for b in range(batch):
        w = W[b].cpu().numpy()
        Q = stack[b, :, :n].cpu().numpy().T
        assert np.all(np.diff(w) >= 0)
        assert relerr(w, np.linalg.eigvalsh(Ks[b])) < 1e-12
        assert relerr((Q * w) @ Q.T, Ks[b]) < 1e-12
        assert relerr(Q.T @ Q, np.eye(n)) < 1e-12


@pytest.mark.parametrize("n,nmat", [(5, 1), (32, 1), (33, 1), (64, 2), (65, 1), (97, 1), (128, 1), (129, 2), (131, 2), (161, 1),
                                    (192, 2), (193, 1), (250, 2), (256, 1)])
def test_cluster_tridiagonalisation_and_backtransform(cuda_lib, n, nmat):
    """M = H T H^T on 8-CTA clusters: T has M's eigenvalues; eigenvectors of T mapped back are eigenvectors of M."""
    import scipy.linalg
    from gpcsd_b200 import _lib as L
    rng = np.random.default_rng(n)
    ld = _ld(n)
    Ms = []
    stack = torch.zeros((nmat, n, ld), dtype=F64, device="cuda")
    for b in range(nmat):
        if b == 0:      # kernel-like matrix with a degenerate tail (the case that matters)
            t = np.arange(n) * 0.7
            dd = t[:, None] - t[None, :]
            K = 0.5 * np.exp(-0.5 * dd ** 2 / 30.0) + 0.2 * np.exp(-np.abs(dd) / 4.0)
        else:
            K = rng.standard_normal((n, n))
            K = K + K.T
        Ms.append(K)
        stack[b, :, :n] = torch.from_numpy(K).cuda()
    d = torch.zeros((nmat, n), dtype=F64, device="cuda")
    e = torch.zeros((nmat, n), dtype=F64, device="cuda")
    V = torch.zeros((nmat, n, ld), dtype=F64, device="cuda")
    tau = torch.zeros((nmat, n), dtype=F64, device="cuda")
    L.call("gpcsd_tridiag", n, nmat, stack.data_ptr(), ld, d.data_ptr(), e.data_ptr(), V.data_ptr(), ld, tau.data_ptr(), _stream())
    XT = torch.zeros((nmat, n, ld), dtype=F64, device="cuda")
    lams = []
    for b in range(nmat):
        dh, eh = d[b].cpu().numpy(), e[b, : n - 1].cpu().numpy()
        lam, X = scipy.linalg.eigh_tridiagonal(dh, eh)
        scale = np.max(np.abs(Ms[b]))
        assert np.max(np.abs(lam - np.linalg.eigvalsh(Ms[b]))) < 1e-13 * scale * n
        XT[b, :, :n] = torch.from_numpy(np.ascontiguousarray(X.T)).cuda()
        lams.append(lam)
    L.call("gpcsd_backtransform", n, nmat, V.data_ptr(), ld, tau.data_ptr(), XT.data_ptr(), ld, _stream())
    for b in range(nmat):
        Q = XT[b, :, :n].cpu().numpy().T
        scale = np.max(np.abs(Ms[b]))
        assert np.max(np.abs(Q.T @ Q - np.eye(n))) < 1e-12
        assert np.max(np.abs(Ms[b] @ Q - Q * lams[b])) < 1e-12 * scale * n


def _tridiag_cases(n, rng):
    t = np.arange(n) * 0.7
    dd = t[:, None] - t[None, :]
    import scipy.linalg
    H = scipy.linalg.hessenberg(3.0 * np.exp(-0.5 * dd ** 2 / 30.0))
    yield np.diag(H).copy(), np.diag(H, -1).copy()                       # SE kernel: numerically degenerate tail
    yield rng.standard_normal(n), rng.standard_normal(n - 1)
    yield np.abs(np.arange(n) - (n - 1) / 2), np.ones(n - 1)             # Wilkinson
    yield 2 * np.ones(n), -np.ones(n - 1)
    yield np.ones(n), np.zeros(n - 1)
    yield np.tile(np.arange(1, 6.0), (n + 4) // 5)[:n], np.where(np.arange(n - 1) % 5 == 4, 1e-9, 1.0)   # glued


@pytest.mark.parametrize("n", [2, 3, 7, 24, 33, 64, 65, 125, 192, 250, 256])
def test_tridiag_eig_divide_and_conquer(cuda_lib, n):
    """Cluster divide-and-conquer eigensolver on tridiagonal matrices vs LAPACK (batched: all cases in one launch)."""
    import scipy.linalg
    from gpcsd_b200 import _lib as L
    rng = np.random.default_rng(n)
    cases = list(_tridiag_cases(n, rng))
    nmat, ld = len(cases), _ld(n)
    d = torch.zeros((nmat, n), dtype=F64, device="cuda")
    e = torch.zeros((nmat, n), dtype=F64, device="cuda")
    for b, (dh, eh) in enumerate(cases):
        d[b] = torch.from_numpy(dh).cuda()
        e[b, : n - 1] = torch.from_numpy(eh).cuda()
    W = torch.zeros((nmat, n), dtype=F64, device="cuda")
    XT = torch.full((nmat, n, ld), 7.0, dtype=F64, device="cuda")
    nws = L.query("gpcsd_tridiag_eig_ws_doubles", n, ld, nmat)
    ws = torch.full((nws,), float("nan"), dtype=F64, device="cuda")      # the workspace may hold anything
    L.call("gpcsd_tridiag_eig", n, nmat, d.data_ptr(), e.data_ptr(), W.data_ptr(), XT.data_ptr(), ld, ws.data_ptr(), nws,
           0, _stream())
    torch.cuda.synchronize()
    for b, (dh, eh) in enumerate(cases):
        T = np.diag(dh) + np.diag(eh, 1) + np.diag(eh, -1)
        lam = scipy.linalg.eigvalsh_tridiagonal(dh, eh) if n > 1 else dh
        sc = np.max(np.abs(T))
        Wh, Q = W[b].cpu().numpy(), XT[b, :, :n].cpu().numpy().T
        assert np.all(np.diff(Wh) >= 0), b
        assert np.max(np.abs(Wh - lam)) <= 5e-14 * sc, b
        assert np.max(np.abs(Q.T @ Q - np.eye(n))) <= 2e-14, b
        assert np.max(np.abs(T @ Q - Q * Wh)) <= 2e-14 * sc, b
        if ld > n:
            assert torch.all(XT[b, :, n:] == 7.0)       # padding untouched


@pytest.mark.parametrize("n,nmat", [(3, 1), (24, 2), (50, 3), (125, 2), (192, 2), (250, 2), (256, 1)])
def test_eigh_dc_full(cuda_lib, n, nmat):
    """gpcsd_eigh_dc (tridiagonalise -> divide and conquer -> back-transform) vs numpy.linalg.eigh on GP covariance factors."""
    from gpcsd_b200 import _lib as L
    rng = np.random.default_rng(n + nmat)
    ld = _ld(n)
    stack = torch.zeros((nmat, n, ld), dtype=F64, device="cuda")
    Ms = []
    for b in range(nmat):
        t = np.arange(n) * (0.5 + b)
        dd = t[:, None] - t[None, :]
        if b == 0:
            K = 0.5 * np.exp(-0.5 * dd ** 2 / 400.0) + 0.2 * np.exp(-np.abs(dd) / 5.0)
        elif b == 1:
            K = np.exp(-0.5 * dd ** 2 / 50.0) + 1e-8 * np.eye(n)
        else:
            K = rng.standard_normal((n, n))
            K = K + K.T
        Ms.append(K)
        stack[b, :, :n] = torch.from_numpy(K).cuda()
    keep = stack.clone()
    QT = torch.zeros((nmat, n, ld), dtype=F64, device="cuda")
    W = torch.zeros((nmat, n), dtype=F64, device="cuda")
    nws = L.query("gpcsd_eigh_dc_ws_doubles", n, ld, nmat)
    ws = torch.full((nws,), float("nan"), dtype=F64, device="cuda")      # the workspace may hold anything
    L.call("gpcsd_eigh_dc", n, nmat, stack.data_ptr(), ld, QT.data_ptr(), ld, W.data_ptr(), ws.data_ptr(), nws, 0, _stream())
    torch.cuda.synchronize()
    assert torch.equal(stack, keep)
    for b in range(nmat):
        lam = np.linalg.eigvalsh(Ms[b])
        sc = np.max(np.abs(lam))
        Wh, Q = W[b].cpu().numpy(), QT[b, :, :n].cpu().numpy().T
        assert np.max(np.abs(Wh - lam)) <= 1e-13 * sc * max(1, n / 16)
        assert np.max(np.abs(Q.T @ Q - np.eye(n))) <= 1e-12
        assert np.max(np.abs(Ms[b] @ Q - Q * Wh)) <= 1e-13 * sc * n


@pytest.mark.parametrize("n", [1, 2, 3, 7, 16, 24, 25, 31, 32])
def test_eigh_small_orders_jacobi(cuda_lib, n):
    """Orders <= 32 take the one-CTA-per-matrix parallel Jacobi kernel: 40 stacked matrices of five families (GP spatial factor
    with condition ~1e10, SE + Matern temporal factor, random symmetric indefinite, repeated eigenvalues, diagonal) vs numpy."""
    from gpcsd_b200 import _lib as L
    rng = np.random.default_rng(100 + n)
    nmat, ld = 40, _ld(n)
    Ms = []
    for b in range(nmat):
        x = np.linspace(0.0, 2300.0, n)
        dd = x[:, None] - x[None, :]
        fam = b % 5
        if fam == 0:
            K = np.exp(-0.5 * dd ** 2 / (200.0 + 30 * b) ** 2) + 1e-8 * np.eye(n)
        elif fam == 1:
            K = 0.5 * np.exp(-0.5 * (dd / 100.0) ** 2 / 400.0) + 0.2 * np.exp(-np.abs(dd / 100.0) / 5.0)
        elif fam == 2:
            K = rng.standard_normal((n, n)); K = K + K.T
        elif fam == 3:
            Qr = np.linalg.qr(rng.standard_normal((n, n)))[0]
            K = (Qr * np.repeat([1.0, 2.0, 2.0, 5.0], n // 4 + 1)[:n]) @ Qr.T
        else:
            K = np.diag(rng.standard_normal(n))
        Ms.append(0.5 * (K + K.T))
    stack = torch.zeros((nmat, n, ld), dtype=F64, device="cuda")
    for b in range(nmat):
        stack[b, :, :n] = torch.from_numpy(Ms[b]).cuda()
    QT = torch.zeros((nmat, n, ld), dtype=F64, device="cuda")
    W = torch.zeros((nmat, n), dtype=F64, device="cuda")
    info = torch.full((nmat,), 7, dtype=torch.int32, device="cuda")
    nws = max(L.query("gpcsd_eigh_dc_ws_doubles", n, ld, nmat), 1)
    ws = torch.zeros(nws, dtype=F64, device="cuda")
    L.call("gpcsd_eigh_dc", n, nmat, stack.data_ptr(), ld, QT.data_ptr(), ld, W.data_ptr(), ws.data_ptr(), nws, info.data_ptr(),
           _stream())
    torch.cuda.synchronize()
    assert torch.all(info == 0)
    for b in range(nmat):
        lam = np.linalg.eigvalsh(Ms[b])
        sc = max(np.max(np.abs(lam)), 1e-300)
        Wh, Q = W[b].cpu().numpy(), QT[b, :, :n].cpu().numpy().T
        assert np.all(np.diff(Wh) >= 0)
        assert np.max(np.abs(Wh - lam)) <= 1e-13 * sc
        assert np.max(np.abs(Q.T @ Q - np.eye(n))) <= 1e-13
        assert np.max(np.abs(Ms[b] @ Q - Q * Wh)) <= 1e-13 * sc * n
        if b % 5 == 0 and n >= 16:
            # graded SPD factor: the small eigenvalues come out to high RELATIVE accuracy (Jacobi's strength)
            pos = lam > 1e-7 * sc
            assert np.max(np.abs(Wh[pos] - lam[pos]) / lam[pos]) < 1e-8


@pytest.mark.parametrize("n", [24, 30, 125, 250])
def test_eigh_dc_nonfinite_input_is_reported(cuda_lib, n):
    """NaN/inf in the matrix (the reference lets them flow, numpy.linalg.eigh then raises): info = 1, NaN outputs, no fault;
    the healthy matrix of the same batch is still solved."""
    from gpcsd_b200 import _lib as L
    ld = _ld(n)
    t = np.arange(n) * 0.5
    K = np.exp(-0.5 * (t[:, None] - t[None, :]) ** 2 / 9.0)
    stack = torch.zeros((3, n, ld), dtype=F64, device="cuda")
    stack[:, :, :n] = torch.from_numpy(K).cuda()
    stack[0, n // 2, n // 3] = float("nan")
    stack[0, n // 3, n // 2] = float("nan")
    stack[2, 1, 1] = float("inf")
    QT = torch.zeros((3, n, ld), dtype=F64, device="cuda")
    W = torch.zeros((3, n), dtype=F64, device="cuda")
    info = torch.full((3,), -1, dtype=torch.int32, device="cuda")
    nws = L.query("gpcsd_eigh_dc_ws_doubles", n, ld, 3)
    ws = torch.zeros(nws, dtype=F64, device="cuda")
    L.call("gpcsd_eigh_dc", n, 3, stack.data_ptr(), ld, QT.data_ptr(), ld, W.data_ptr(), ws.data_ptr(), nws, info.data_ptr(),
           _stream())
    torch.cuda.synchronize()
    assert info.cpu().tolist() == [1, 0, 1]
    assert torch.isnan(W[0]).all() and torch.isnan(W[2]).all()
    lam = np.linalg.eigvalsh(K)
    assert np.max(np.abs(W[1].cpu().numpy() - lam)) <= 1e-13 * lam.max() * max(1, n / 16)


@pytest.mark.parametrize("n,nblk,rowlen", [(6, 2, 8), (7, 3, 16), (50, 4, 24), (501, 2, 40)])
def test_centro_fold(cuda_lib, n, nblk, rowlen):
    """Folded time basis: sums / differences of mirrored time rows, middle row kept for odd n."""
    from gpcsd_b200 import _lib as L
    rng = np.random.default_rng(n)
    X = rng.standard_normal((nblk, n, rowlen))
    Xd = torch.from_numpy(X).cuda()
    Xf = torch.full_like(Xd, 7.0)
    L.call("gpcsd_centro_fold", nblk, n, rowlen, Xd.data_ptr(), Xf.data_ptr(), _stream())
    m, ms = n // 2, n - n // 2
    ref = np.empty_like(X)
    ref[:, :m] = (X[:, :m] + X[:, ::-1][:, :m]) / np.sqrt(2)
    if n & 1:
        ref[:, m] = X[:, m]
    ref[:, ms:] = (X[:, :m] - X[:, ::-1][:, :m]) / np.sqrt(2)
    assert relerr(Xf.cpu().numpy(), ref) < 1e-15
    back = torch.full_like(Xd, 7.0)
    L.call("gpcsd_centro_unfold", nblk, n, rowlen, Xf.data_ptr(), back.data_ptr(), _stream())
    assert relerr(back.cpu().numpy(), X) < 1e-15        # the fold is orthogonal: unfold is its exact inverse up to rounding


def test_pairsym_fold(cuda_lib):
    """Channel fold under an index pairing: sums / differences of paired rows."""
    from gpcsd_b200 import _lib as L
    rng = np.random.default_rng(5)
    n, rowlen = 12, 40
    perm = rng.permutation(n)
    ra, rb = perm[: n // 2].astype(np.int32), perm[n // 2:].astype(np.int32)
    X = rng.standard_normal((n, rowlen))
    Xd = torch.from_numpy(X).cuda()
    Xf = torch.zeros_like(Xd)
    rad, rbd = torch.from_numpy(ra).cuda(), torch.from_numpy(rb).cuda()
    L.call("gpcsd_pairsym_fold", n, rad.data_ptr(), rbd.data_ptr(), rowlen, Xd.data_ptr(), Xf.data_ptr(), _stream())
    ref = np.concatenate([(X[ra] + X[rb]) / np.sqrt(2), (X[ra] - X[rb]) / np.sqrt(2)])
    assert relerr(Xf.cpu().numpy(), ref) < 1e-15
