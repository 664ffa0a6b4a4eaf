"""numpy-in / numpy-out wrappers over the C ABI for the covariance helper API.

The covariance classes of the reference return host matrices (covariances.py); callers outside the
engine (scripts, ``sample_prior``) expect the same.  Each helper uploads its small inputs, runs the CUDA
kernels of libgpcsd_b200.so and downloads the result -- no numpy arithmetic and no CPU fallback.
"""
import numpy as np
import torch

from . import _lib as L

F64 = torch.float64


def _require_cuda():
    if not torch.cuda.is_available():
        raise L.GpcsdLibraryError("gpcsd_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    L.load()


def _even(n):
    return (int(n) + 1) // 2 * 2


def _dev(a):
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).cuda()


def _mat(rows, cols):
    return torch.zeros((int(rows), _even(cols)), dtype=F64, device="cuda")


def _upload_mat(a):
    a = np.atleast_2d(np.asarray(a, dtype=np.float64))
    m = _mat(a.shape[0], a.shape[1])
    m[:, : a.shape[1]] = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return m


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _pad_even(nodes, weights):
    nodes, weights = np.asarray(nodes, dtype=np.float64).reshape(-1), np.asarray(weights, dtype=np.float64).reshape(-1)
    if len(nodes) % 2:
        nodes, weights = np.append(nodes, nodes[-1]), np.append(weights, 0.0)
    return nodes, weights


def gemm(transB, M, N, K, A, lda, B, ldb, C, ldc):
    L.call("gpcsd_dgemm", int(transB), int(M), int(N), int(K), A.data_ptr(), lda, 0, B.data_ptr(), ldb, 0,
           C.data_ptr(), ldc, 0, 1, _stream())


def se_matrix(a, b, ell, deriv=0):
    """exp(-0.5 ((a_i - b_j)/ell)^2)  (covariances.py:56, 67, 89)."""
    _require_cuda()
    a, b = np.asarray(a, dtype=np.float64).reshape(-1), np.asarray(b, dtype=np.float64).reshape(-1)
    ad, bd = _dev(a), _dev(b)
    out = _mat(len(a), len(b))
    L.call("gpcsd_se_matrix", len(a), ad.data_ptr(), len(b), bd.data_ptr(), float(ell), 1.0, int(deriv), out.data_ptr(),
           out.shape[1], _stream())
    return out[:, : len(b)].cpu().numpy()


def kt(kind, ell, sigma2, t, tprime):
    """sigma2 * f((t - tprime^T)/ell): compute_Kt (covariances.py:257-271, 291-305)."""
    _require_cuda()
    t = np.asarray(t, dtype=np.float64).reshape(-1)
    tp = np.asarray(tprime, dtype=np.float64).reshape(-1)
    td, tpd = _dev(t), _dev(tp)
    out = _mat(len(t), len(tp))
    L.call("gpcsd_kt_build", len(t), td.data_ptr(), len(tp), tpd.data_ptr(), 1, L.c_int_array([kind]),
           L.c_double_array([ell]), L.c_double_array([sigma2]), out.data_ptr(), out.shape[1], _stream())
    return out[:, : len(tp)].cpu().numpy()


def _weights_1d(pts, gx, gw, R):
    pd = _dev(np.asarray(pts, dtype=np.float64).reshape(-1))
    G = len(gx)
    A = torch.zeros((pd.shape[0], G), dtype=F64, device="cuda")
    gxd, gwd = _dev(gx), _dev(gw)
    L.call("gpcsd_fwd_weights_1d", pd.shape[0], pd.data_ptr(), G, gxd.data_ptr(), gwd.data_ptr(), float(R),
           A.data_ptr(), None, G, _stream())
    return A


def kphi_1d(x, gl_x, gl_w, R, ell, xp=None):
    """(A Kg) A'^T  -- compKphi_1d (covariances.py:74-96)."""
    _require_cuda()
    gx, gw = _pad_even(gl_x, gl_w)
    G = len(gx)
    A = _weights_1d(x, gx, gw, R)
    Ap = A if xp is None else _weights_1d(xp, gx, gw, R)
    gxd = _dev(gx)
    Kg = torch.zeros((G, G), dtype=F64, device="cuda")
    L.call("gpcsd_se_matrix", G, gxd.data_ptr(), G, gxd.data_ptr(), float(ell), 1.0, 0, Kg.data_ptr(), G, _stream())
    nx, nxp = A.shape[0], Ap.shape[0]
    U = torch.zeros((nx, G), dtype=F64, device="cuda")
    gemm(0, nx, G, G, A, G, Kg, G, U, G)
    out = _mat(nx, nxp)
    gemm(1, nx, nxp, G, U, G, Ap, G, out, out.shape[1])
    return out[:, :nxp].cpu().numpy()


def kphig_1d(x, gl_x, gl_w, z, R, ell):
    """A exp(-0.5((gl - z)/ell)^2)^T -- compKphig_1d (covariances.py:58-72); (nx, nz)."""
    _require_cuda()
    gx, gw = _pad_even(gl_x, gl_w)
    G = len(gx)
    A = _weights_1d(x, gx, gw, R)
    zd, gxd = _dev(np.asarray(z, dtype=np.float64).reshape(-1)), _dev(gx)
    nz, nx = zd.shape[0], A.shape[0]
    KgzT = torch.zeros((nz, G), dtype=F64, device="cuda")
    L.call("gpcsd_se_matrix", nz, zd.data_ptr(), G, gxd.data_ptr(), float(ell), 1.0, 0, KgzT.data_ptr(), G, _stream())
    out = _mat(nx, nz)
    gemm(1, nx, nz, G, A, G, KgzT, G, out, out.shape[1])
    return out[:, :nz].cpu().numpy()


class Quad2D:
    """Device copy of the 2-D product quadrature (covariances.py:114-131) with the x2 axis padded to even."""

    def __init__(self, gl_x1, gl_w1, gl_x2, gl_w2):
        _require_cuda()
        g2, w2 = _pad_even(gl_x2, gl_w2)
        self.G1, self.G2 = len(np.asarray(gl_x1).reshape(-1)), len(g2)
        self.G = self.G1 * self.G2
        self.g1, self.w1, self.g2, self.w2 = _dev(gl_x1), _dev(gl_w1), _dev(g2), _dev(w2)

    def weights(self, pts, R, eps):
        pd = _dev(np.asarray(pts, dtype=np.float64).reshape(-1, 2))
        A = torch.zeros((pd.shape[0], self.G), dtype=F64, device="cuda")
        L.call("gpcsd_fwd_weights_2d", pd.shape[0], pd.data_ptr(), self.G1, self.G2, self.g1.data_ptr(), self.w1.data_ptr(),
               self.g2.data_ptr(), self.w2.data_ptr(), float(R), float(eps), A.data_ptr(), None, self.G, _stream())
        return A

    def apply_kernel(self, X, ell1, ell2):
        """X (n x G) times the SE kernel on the product grid = Kronecker product of two 1-D SE factors."""
        n = X.shape[0]
        K1, K2 = _mat(self.G1, self.G1), _mat(self.G2, self.G2)
        L.call("gpcsd_se_matrix", self.G1, self.g1.data_ptr(), self.G1, self.g1.data_ptr(), float(ell1), 1.0, 0,
               K1.data_ptr(), K1.shape[1], _stream())
        L.call("gpcsd_se_matrix", self.G2, self.g2.data_ptr(), self.G2, self.g2.data_ptr(), float(ell2), 1.0, 0,
               K2.data_ptr(), K2.shape[1], _stream())
        T = torch.zeros((n, self.G), dtype=F64, device="cuda")
        U = torch.zeros((n, self.G), dtype=F64, device="cuda")
        gemm(0, n * self.G1, self.G2, self.G2, X, self.G2, K2, K2.shape[1], T, self.G2)
        L.call("gpcsd_dgemm", 0, self.G1, self.G2, self.G1, K1.data_ptr(), K1.shape[1], 0, T.data_ptr(), self.G2, self.G,
               U.data_ptr(), self.G2, self.G, n, _stream())
        return U


def kphi_2d(quad, x, R, eps, ell1, ell2, xp=None):
    """compKphi_2d (covariances.py:204-232)."""
    A = quad.weights(x, R, eps)
    Ap = A if xp is None else quad.weights(xp, R, eps)
    U = quad.apply_kernel(A, ell1, ell2)
    nx, nxp = A.shape[0], Ap.shape[0]
    out = _mat(nx, nxp)
    gemm(1, nx, nxp, quad.G, U, quad.G, Ap, quad.G, out, out.shape[1])
    return out[:, :nxp].cpu().numpy()


def kphig_2d(quad, x, z, R, eps, ell1, ell2):
    """compKphig_2d (covariances.py:188-202); (nx, nz)."""
    A = quad.weights(x, R, eps)
    zd = _dev(np.asarray(z, dtype=np.float64).reshape(-1, 2))
    nz, nx = zd.shape[0], A.shape[0]
    KgzT = torch.zeros((nz, quad.G), dtype=F64, device="cuda")
    L.call("gpcsd_se_grid_to_pts", quad.G1, quad.G2, quad.g1.data_ptr(), quad.g2.data_ptr(), nz, zd.data_ptr(),
           float(ell1), float(ell2), KgzT.data_ptr(), quad.G, _stream())
    out = _mat(nx, nz)
    gemm(1, nx, nz, quad.G, A, quad.G, KgzT, quad.G, out, out.shape[1])
    return out[:, :nz].cpu().numpy()


def eigh(K):
    """Ascending eigenvalues and eigenvectors (columns) of a symmetric matrix: the in-house cluster solver
    (gpcsd_eigh_dc) for orders 3..256, cuSOLVER syevd (gpcsd_eigh) otherwise."""
    _require_cuda()
    K = np.asarray(K, dtype=np.float64)
    n = K.shape[0]
    Kd = _upload_mat(K)
    ld = Kd.shape[1]
    QT = torch.zeros((n, ld), dtype=F64, device="cuda")
    W = torch.zeros(n, dtype=F64, device="cuda")
    if 3 <= n <= 256:
        nws = L.query("gpcsd_eigh_dc_ws_doubles", n, ld, 1)
        ws = torch.empty(nws, dtype=F64, device="cuda")
        info = torch.zeros(1, dtype=torch.int32, device="cuda")
        L.call("gpcsd_eigh_dc", n, 1, Kd.data_ptr(), ld, QT.data_ptr(), ld, W.data_ptr(), ws.data_ptr(), nws, info.data_ptr(),
               _stream())
        if int(info.item()) != 0:
            raise np.linalg.LinAlgError("Eigenvalues did not converge")
        return W.cpu().numpy(), QT[:, :n].cpu().numpy().T.copy()
    nws = L.query("gpcsd_eigh_ws_doubles", n, ld)
    ws = torch.zeros(max(nws, 1), dtype=F64, device="cuda")
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    L.call("gpcsd_eigh", n, Kd.data_ptr(), ld, QT.data_ptr(), ld, W.data_ptr(), ws.data_ptr(), nws, info.data_ptr(), _stream())
    if int(info.item()) != 0:
        raise np.linalg.LinAlgError("Eigenvalues did not converge")
    return W.cpu().numpy(), QT[:, :n].cpu().numpy().T.copy()


def comp_eig_D(Ks, Kt, sig2n):
    """utility_functions.py:44-64 on the GPU: returns (evec_s, evec_t, Dvec)."""
    ls, Qs = eigh(Ks)
    lt, Qt = eigh(Kt)
    nx, nt = len(ls), len(lt)
    s = np.atleast_1d(np.asarray(sig2n, dtype=np.float64))
    lsd, ltd, sd = _dev(ls), _dev(lt), _dev(s)
    ldrd = _even(nt)
    rD = torch.zeros((nx, ldrd), dtype=F64, device="cuda")
    sums, rowA, rowC, rowL = (torch.zeros(k, dtype=F64, device="cuda") for k in (2, nx, nx, nx))
    colB = torch.zeros(nt, dtype=F64, device="cuda")
    L.call("gpcsd_eig_D", nx, nt, lsd.data_ptr(), ltd.data_ptr(), sd.data_ptr(), len(s), rD.data_ptr(), ldrd,
           sums.data_ptr(), rowA.data_ptr(), rowC.data_ptr(), rowL.data_ptr(), colB.data_ptr(), _stream())
    Dvec = (1.0 / rD[:, :nt]).reshape(-1).cpu().numpy()
    return Qs, Qt, Dvec


def sandwich(Ls, X, Lt):
    """out[:, :, r] = Ls X[:, :, r] Lt^T for all r: two DMMA GEMMs on the trial-fastest layout
    (the per-trial product of sample_prior, gpcsd1d.py:307-308 / gpcsd2d.py:355-359); host in, host out."""
    _require_cuda()
    X = np.ascontiguousarray(np.atleast_3d(np.asarray(X, dtype=np.float64)))
    nx, nt, N = X.shape
    ldn = (N + 7) // 8 * 8
    Xd = torch.zeros((nx, nt, ldn), dtype=F64, device="cuda")
    Xd[:, :, :N] = torch.from_numpy(X).cuda()
    Lsd, Ltd = _upload_mat(Ls), _upload_mat(Lt)
    W = torch.zeros_like(Xd)
    out = torch.zeros_like(Xd)
    L.call("gpcsd_dgemm", 0, nx, nt * ldn, nx, Lsd.data_ptr(), Lsd.shape[1], 0, Xd.data_ptr(), nt * ldn, 0,
           W.data_ptr(), nt * ldn, 0, 1, _stream())
    L.call("gpcsd_dgemm", 0, nt, ldn, nt, Ltd.data_ptr(), Ltd.shape[1], 0, W.data_ptr(), ldn, nt * ldn,
           out.data_ptr(), ldn, nt * ldn, nx, _stream())
    return np.ascontiguousarray(out[:, :, :N].cpu().numpy())


# ------------------------------------------------------------------------------------------------------------------
# forward-model operators (forward_models.py:20-39, 57-81): weight-matrix kernel + ONE DMMA GEMM
# ------------------------------------------------------------------------------------------------------------------
def fwd_apply_1d(arr, x, z, R, varsigma=1.0):
    """R/(2 varsigma) * trapz_x b_fwd_1d(z_i - x, R) arr[:, t] for all (i, t): (nz, nt)."""
    _require_cuda()
    arr = np.asarray(arr, dtype=np.float64)
    if arr.ndim == 1:
        arr = arr[:, None]
    xd, zd = _dev(np.asarray(x, dtype=np.float64).reshape(-1)), _dev(np.asarray(z, dtype=np.float64).reshape(-1))
    nx, nz, nt = xd.shape[0], zd.shape[0], arr.shape[1]
    if arr.shape[0] != nx:
        raise ValueError("operands could not be broadcast together with shapes (%d,1) (%d,1)" % (nx, arr.shape[0]))
    W = _mat(nz, nx)
    L.call("gpcsd_fwd_operator_1d", nz, zd.data_ptr(), nx, xd.data_ptr(), float(R), float(R) / (2.0 * float(varsigma)),
           W.data_ptr(), W.shape[1], _stream())
    Ad = _upload_mat(arr)
    out = _mat(nz, nt)
    gemm(0, nz, nt, nx, W, W.shape[1], Ad, Ad.shape[1], out, out.shape[1])
    return out[:, :nt].cpu().numpy()


def fwd_apply_2d(arr, x1, x2, z, R, eps):
    """Double trapezoid of b_fwd_2d * arr over the grid x1 x x2 for every z and time point: (nz, nt)."""
    _require_cuda()
    arr = np.asarray(arr, dtype=np.float64)
    x1d, x2d = _dev(np.asarray(x1, dtype=np.float64).reshape(-1)), _dev(np.asarray(x2, dtype=np.float64).reshape(-1))
    zd = _dev(np.asarray(z, dtype=np.float64).reshape(-1, 2))
    n1, n2, nz = x1d.shape[0], x2d.shape[0], zd.shape[0]
    arr = arr.reshape(n1 * n2, -1)
    nt = arr.shape[1]
    W = _mat(nz, n1 * n2)
    L.call("gpcsd_fwd_operator_2d", nz, zd.data_ptr(), n1, x1d.data_ptr(), n2, x2d.data_ptr(), float(R), float(eps),
           W.data_ptr(), W.shape[1], _stream())
    Ad = _upload_mat(arr)
    out = _mat(nz, nt)
    gemm(0, nz, nt, n1 * n2, W, W.shape[1], Ad, Ad.shape[1], out, out.shape[1])
    return out[:, :nt].cpu().numpy()


# ------------------------------------------------------------------------------------------------------------------
# sample_prior on the device (gpcsd1d.py:295-309 / gpcsd2d.py:336-360): Cholesky, Philox normals, two GEMMs
# ------------------------------------------------------------------------------------------------------------------
def cholesky_device(K):
    """Lower Cholesky factor of a host (or device [n][ld]) symmetric matrix, left on the device as [n][ld]."""
    _require_cuda()
    Ld = _upload_mat(K) if not isinstance(K, torch.Tensor) else K.clone()
    n = Ld.shape[0]
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    L.call("gpcsd_cholesky", n, Ld.data_ptr(), Ld.shape[1], info.data_ptr(), _stream())
    if int(info.item()) != 0:
        raise np.linalg.LinAlgError("Matrix is not positive definite")
    return Ld


def cholesky(K):
    """np.linalg.cholesky on the GPU (host in, host out)."""
    K = np.asarray(K, dtype=np.float64)
    return cholesky_device(K)[:, : K.shape[0]].cpu().numpy()


def randn_device(nrows, ncols, seed, stream_id=0, ld=None, sd=1.0, out=None, accumulate=False):
    """[nrows][ld] device array whose first ncols columns are N(0, sd^2) draws of the Philox4x32-10 stream (seed, stream_id)."""
    _require_cuda()
    ld = int(ld or ncols)
    if out is None:
        out = torch.zeros((int(nrows), ld), dtype=F64, device="cuda")
    L.call("gpcsd_randn", int(nrows), int(ncols), ld, int(seed) & (2 ** 64 - 1), int(stream_id) & 0xFFFFFFFF, float(sd),
           1 if accumulate else 0, out.data_ptr(), _stream())
    return out


def randn(shape, seed, stream_id=0):
    """Host array of standard normals from the device generator; flat C-order element i is normal (i & 1) of counter i >> 1."""
    n = int(np.prod(shape))
    return randn_device(1, n, seed, stream_id)[0].cpu().numpy().reshape(shape)


def sample_gp_device(Ls, Lt, ntrials, seed, noise_sd=0.0, rand=None):
    """Device array [nx][nt][ldn] with out[:, :, r] = Ls Z_r Lt^T (+ noise_sd * E_r): Z from the Philox stream (seed, 0) -- or
    the caller's `rand` (nx, nt, ntrials) host array --, E from stream (seed, 1) added in place without materialising it.
    Ls, Lt: device [n][ld] lower factors (cholesky_device) or host matrices."""
    _require_cuda()
    Lsd = Ls if isinstance(Ls, torch.Tensor) else _upload_mat(Ls)
    Ltd = Lt if isinstance(Lt, torch.Tensor) else _upload_mat(Lt)
    nx, nt, N = Lsd.shape[0], Ltd.shape[0], int(ntrials)
    ldn = (N + 7) // 8 * 8
    if rand is None:
        Z = randn_device(nx * nt, N, seed, 0, ld=ldn).view(nx, nt, ldn)
    else:
        Z = torch.zeros((nx, nt, ldn), dtype=F64, device="cuda")
        Z[:, :, :N] = torch.from_numpy(np.ascontiguousarray(rand, dtype=np.float64)).cuda()
    W = torch.empty_like(Z)
    L.call("gpcsd_dgemm", 0, nx, nt * ldn, nx, Lsd.data_ptr(), Lsd.shape[1], 0, Z.data_ptr(), nt * ldn, 0,
           W.data_ptr(), nt * ldn, 0, 1, _stream())
    out = Z                                             # reuse: Z is dead after the first product
    L.call("gpcsd_dgemm", 0, nt, ldn, nt, Ltd.data_ptr(), Ltd.shape[1], 0, W.data_ptr(), ldn, nt * ldn,
           out.data_ptr(), ldn, nt * ldn, nx, _stream())
    if noise_sd:
        randn_device(nx * nt, N, seed, 1, ld=ldn, sd=noise_sd, out=out.view(nx * nt, ldn), accumulate=True)
    return out, N
