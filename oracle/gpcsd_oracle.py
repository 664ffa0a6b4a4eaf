"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's GPCSD hot path.

This file is the CPU oracle the CUDA path is checked against.  It restates, in plain numpy, the
arithmetic of natalieklein/gpcsd's ``loglik`` / ``obj_fun`` / ``predict`` and of the covariance and
forward-model helpers they call.  Every function cites the reference ``file:line`` it follows
(paths relative to ``/root/reference/src/gpcsd``).

Pinning status
--------------
* Forward path (covariances, ``comp_eig_D``, ``loglik``, ``predict``): the reference has no tests and no
  golden vectors.  The oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, imported here through
  ``oracle/ref_shim.py`` on explicit seeded inputs; the resulting vectors are committed under
  ``tests/golden/`` together with the generating script ``oracle/make_golden.py``.
* Gradient path: the reference's gradient is ``autograd.grad(obj_fun)`` (HIPS autograd, an un-pinned,
  un-vendored PyPI dependency -- ``setup.py:30`` -- absent from this image).  No reference test pins
  it.  **Gradient parity is therefore UNPINNED by the reference**; it is anchored on (i) the published
  algorithm of reverse-mode AD = the exact derivative of the reference's ``obj_fun``, restated twice
  independently (closed form below; torch-float64 autograd through ``torch.linalg.eigh`` in
  ``oracle/oracle_torch.py``) and (ii) 4th-order central finite differences of the REFERENCE's own
  ``obj_fun``/``loglik`` (golden fixtures carry those FD gradients).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this.
"""
from dataclasses import dataclass, field
from typing import List, Tuple, Union

import numpy as np
import scipy.special

JITTER_1D = 1e-8  # gpcsd1d.py:17
JITTER_2D = 1e-7  # gpcsd2d.py:16


# --------------------------------------------------------------------------------------------
# forward-model weights  (forward_models.py)
# --------------------------------------------------------------------------------------------
def b_fwd_1d(r, R):
    """forward_models.py:9-17 -- sqrt((r/R)^2+1) - sqrt((r/R)^2)."""
    q = np.square(r / R)
    return np.sqrt(q + 1.0) - np.sqrt(q)


def b_fwd_2d(w, R, eps):
    """forward_models.py:42-54 with the distance ``w`` already formed (covariances.py:131)."""
    Re = R + eps
    return np.log(Re + np.sqrt(Re ** 2 + w ** 2)) - np.log(eps + np.sqrt(eps ** 2 + w ** 2))


def fwd_model_1d(arr, x, z, R, varsigma=1.0):
    """forward_models.py:20-39 -- trapezoid CSD->LFP operator, written as one weight-matrix product.

    res[i, t] = trapz(b(z_i - x, R) * arr[:, t], x)  ==  sum_k W[i, k] arr[k, t] with trapezoid
    weights folded into W."""
    xs = np.squeeze(x)
    dx = np.diff(xs)
    tw = np.zeros_like(xs)
    tw[:-1] += 0.5 * dx
    tw[1:] += 0.5 * dx
    W = b_fwd_1d(np.asarray(z).reshape(-1, 1) - xs[None, :], R) * tw[None, :]
    return R / (2.0 * varsigma) * (W @ arr)


# --------------------------------------------------------------------------------------------
# quadrature (covariances.py:12-27, 99-131)
# --------------------------------------------------------------------------------------------
def gauss_legendre(a, b, n):
    """covariances.py:22-27 -- GL nodes/weights mapped from [-1,1] to [a,b]."""
    u, w = scipy.special.roots_legendre(n)
    return 0.5 * (u + 1.0) * (b - a) + a, 0.5 * (b - a) * w


@dataclass
class Spatial1D:
    """State of GPCSD1DSpatialCov (covariances.py:12-27)."""
    x: np.ndarray  # (nx, 1)
    a: float
    b: float
    ngl: int = 100
    gl_x: np.ndarray = field(init=False)
    gl_w: np.ndarray = field(init=False)

    def __post_init__(self):
        self.x = np.asarray(self.x, dtype=np.float64).reshape(-1, 1)
        self.gl_x, self.gl_w = gauss_legendre(self.a, self.b, self.ngl)


@dataclass
class Spatial2D:
    """State of GPCSD2DSpatialCov (covariances.py:99-131); product grid is x1-major / x2-minor
    (utility_functions.py:22)."""
    x: np.ndarray  # (nx, 2)
    a1: float
    b1: float
    a2: float
    b2: float
    ngl1: int = 20
    ngl2: int = 60

    def __post_init__(self):
        self.x = np.asarray(self.x, dtype=np.float64)
        self.gl_x1, self.gl_w1 = gauss_legendre(self.a1, self.b1, self.ngl1)
        self.gl_x2, self.gl_w2 = gauss_legendre(self.a2, self.b2, self.ngl2)
        self.grid1 = np.repeat(self.gl_x1, self.ngl2)      # first column of gl_x_grid
        self.grid2 = np.tile(self.gl_x2, self.ngl1)        # second column
        self.w_prod = np.repeat(self.gl_w1, self.ngl2) * np.tile(self.gl_w2, self.ngl1)  # cov:126

    def delta_w(self, pts):
        """cov:127-131 -- Euclidean distance from each site in ``pts`` to each quadrature node."""
        d1 = self.grid1[None, :] - pts[:, 0][:, None]
        d2 = self.grid2[None, :] - pts[:, 1][:, None]
        return np.sqrt(np.square(d1) + np.square(d2))


# --------------------------------------------------------------------------------------------
# spatial covariances
# --------------------------------------------------------------------------------------------
def compKphi_1d(sp: Spatial1D, R, ell, xp=None):
    """covariances.py:74-96 -- LFP-LFP spatial covariance (A Kg) A'^T."""
    xp = sp.x if xp is None else np.asarray(xp, dtype=np.float64).reshape(-1, 1)
    g = sp.gl_x[None, :]
    A = sp.gl_w[None, :] * b_fwd_1d(g - sp.x, R)                       # cov:86-88
    Kg = np.exp(-0.5 * np.square((g.T - g) / ell))                     # cov:89
    Ap = sp.gl_w[None, :] * b_fwd_1d(g - xp, R)                        # cov:92-94
    return np.dot(np.dot(A, Kg), Ap.T)                                 # cov:90,95


def compKphig_1d(sp: Spatial1D, z, R, ell):
    """covariances.py:58-72 -- LFP-CSD cross covariance, (nx, nz)."""
    z = np.asarray(z, dtype=np.float64).reshape(-1, 1)
    g = sp.gl_x[None, :]
    Kgz = np.exp(-0.5 * np.square((g - z) / ell)).T                    # cov:67  (ngl, nz)
    A = sp.gl_w[None, :] * b_fwd_1d(g - sp.x, R)                       # cov:68-70
    return np.dot(A, Kgz)


def compute_Ks_1d(x, ell):
    """covariances.py:50-56 -- CSD-CSD SE kernel at sites x."""
    x = np.asarray(x, dtype=np.float64).reshape(-1, 1)
    return np.exp(-0.5 * np.square(x - x.T) / np.square(ell))


def compKphi_2d(sp: Spatial2D, R, eps, ell1, ell2, xp=None):
    """covariances.py:204-232."""
    sq1 = np.square(sp.grid1[:, None] - sp.grid1[None, :])             # cov:129
    sq2 = np.square(sp.grid2[:, None] - sp.grid2[None, :])             # cov:130
    Kg = np.exp(-0.5 * sq1 / ell1 ** 2) * np.exp(-0.5 * sq2 / ell2 ** 2)  # cov:216
    A = sp.w_prod[None, :] * b_fwd_2d(sp.delta_w(sp.x), R, eps)        # cov:220-221
    U = np.matmul(A, Kg)                                               # cov:223
    if xp is not None:
        A = sp.w_prod[None, :] * b_fwd_2d(sp.delta_w(np.asarray(xp, dtype=np.float64)), R, eps)  # cov:226-229
    return np.matmul(U, A.T)                                           # cov:231


def compKphig_2d(sp: Spatial2D, z, R, eps, ell1, ell2):
    """covariances.py:188-202 -- (nx, nz)."""
    z = np.asarray(z, dtype=np.float64)
    Kgz = (np.exp(-0.5 * np.square((sp.grid1[:, None] - z[:, 0][None, :]) / ell1))
           * np.exp(-0.5 * np.square((sp.grid2[:, None] - z[:, 1][None, :]) / ell2)))  # cov:198
    A = sp.w_prod[None, :] * b_fwd_2d(sp.delta_w(sp.x), R, eps)        # cov:199-200
    return np.dot(A, Kgz)


def compute_Ks_2d(x, ell1, ell2):
    """covariances.py:177-186."""
    x = np.asarray(x, dtype=np.float64)
    x1 = x[:, 0][:, None]
    x2 = x[:, 1][:, None]
    return np.exp(-0.5 * np.square((x1 - x1.T) / ell1)) * np.exp(-0.5 * np.square((x2 - x2.T) / ell2))


# --------------------------------------------------------------------------------------------
# temporal covariances (covariances.py:257-271, 291-305)
# --------------------------------------------------------------------------------------------
KIND_SE = 0
KIND_MATERN = 1


def compute_Kt(kind, ell, sigma2, t, tprime=None):
    """SE: cov:269-270;  Matern-1/2: cov:303-304.  ``t`` (n,1) rows, ``tprime`` (m,1) columns."""
    t = np.asarray(t, dtype=np.float64).reshape(-1, 1)
    tprime = t if tprime is None else np.asarray(tprime, dtype=np.float64).reshape(-1, 1)
    dist = t - tprime.T
    if kind == KIND_SE:
        return sigma2 * np.exp(-0.5 * np.square(dist) / np.square(ell))
    if kind == KIND_MATERN:
        return sigma2 * np.exp(-np.sqrt(np.square(dist)) / ell)
    raise ValueError("unknown temporal kernel kind %r" % (kind,))


# --------------------------------------------------------------------------------------------
# model specification shared by 1-D and 2-D
# --------------------------------------------------------------------------------------------
@dataclass
class Model:
    """Everything ``loglik``/``predict`` read from a GPCSD1D / GPCSD2D object."""
    dim: int                                    # 1 or 2
    spatial: Union[Spatial1D, Spatial2D]
    t: np.ndarray                               # (nt, 1)
    R: float
    ells: Tuple[float, ...]                     # (ell,) or (ell1, ell2)
    temporal: List[Tuple[int, float, float]]    # [(kind, ell, sigma2), ...] in list order
    sig2n: Union[float, np.ndarray]             # scalar, or vector indexed by ASCENDING SPATIAL EIGENVALUE (util:54-57)
    eps: float = 0.0                            # 2-D only (gpcsd2d.py:69-71)

    @property
    def jitter(self):
        return JITTER_1D if self.dim == 1 else JITTER_2D

    def Ks(self, xp=None, jitter=False):
        if self.dim == 1:
            K = compKphi_1d(self.spatial, self.R, self.ells[0], xp=xp)
        else:
            K = compKphi_2d(self.spatial, self.R, self.eps, self.ells[0], self.ells[1], xp=xp)
        if jitter:
            K = K + self.jitter * np.eye(K.shape[0])                   # 1d:117 / 2d:140
        return K

    def Kphig(self, z):
        if self.dim == 1:
            return compKphig_1d(self.spatial, z, self.R, self.ells[0])
        return compKphig_2d(self.spatial, z, self.R, self.eps, self.ells[0], self.ells[1])

    def Kt(self):
        nt = len(self.t)
        K = np.zeros((nt, nt))
        for kind, ell, s2 in self.temporal:                            # 1d:118-120
            K = K + compute_Kt(kind, ell, s2, self.t)
        return K


def eigh_driver(driver):
    """np.linalg.eigh stand-in using another LAPACK driver ('ev', 'evd', 'evr', 'evx'): used by tests to
    measure the reference formula's own solver-to-solver spread."""
    import scipy.linalg
    return lambda K: scipy.linalg.eigh(K, driver=driver)


def comp_eig_D(Ks, Kt, sig2n, eigh=np.linalg.eigh):
    """utility_functions.py:44-64 -- two eigh (ascending), D[i*nt+j] = ls_i*lt_j + sig2n(_i)."""
    nx, nt = Ks.shape[0], Kt.shape[0]
    if np.isscalar(sig2n) or np.ndim(sig2n) == 0:
        nvec = float(sig2n) * np.ones(nx * nt)
    else:
        nvec = np.repeat(np.asarray(sig2n, dtype=np.float64), nt)      # util:57: by spatial EIGEN index
    lt, Qt = eigh(Kt)
    ls, Qs = eigh(Ks)
    D = np.repeat(ls, nt) * np.tile(lt, nx) + nvec
    return Qs, Qt, D, ls, lt


def loglik(model: Model, lfp, eigh=np.linalg.eigh):
    """gpcsd1d.py:113-128 / gpcsd2d.py:136-151 -- literal restatement, Python trial loop included."""
    lfp = np.atleast_3d(lfp)
    nx, nt, ntrials = lfp.shape
    Ks = model.Ks(jitter=True)
    Kt = model.Kt()
    Qs, Qt, D, _, _ = comp_eig_D(Ks, Kt, model.sig2n, eigh)
    logdet = -0.5 * ntrials * np.sum(np.log(D))
    quad = 0.0
    for r in range(ntrials):
        alpha = np.reshape(np.dot(np.dot(Qs.T, lfp[:, :, r]), Qt), nx * nt)
        quad = quad + np.sum(np.square(alpha) / D)
    return float(logdet - 0.5 * quad)


def loglik_from_factors(lfp, Qs, ls, Qt, lt, sig2n):
    """Same sum as ``loglik`` given the eigen-factors (kernel-level oracle: identical Qs,ls,Qt,lt)."""
    nx, nt, N = lfp.shape
    s = np.broadcast_to(np.asarray(sig2n, dtype=np.float64), (nx,)) if np.ndim(sig2n) else np.full(nx, float(sig2n))
    D = ls[:, None] * lt[None, :] + s[:, None]
    A = np.einsum("ia,ijr,jb->abr", Qs, lfp, Qt, optimize=True)
    return float(-0.5 * N * np.sum(np.log(D)) - 0.5 * np.sum(A * A / D[:, :, None]))


# --------------------------------------------------------------------------------------------
# priors (priors.py:23-28, 46-51) and the fit objective (gpcsd1d.py:153-191 / gpcsd2d.py:177-221)
# --------------------------------------------------------------------------------------------
def invgamma_lpdf(x, alpha, beta):
    return -np.inf if x <= 0 else -(alpha + 1.0) * np.log(x) - beta / x


def halfnormal_lpdf(x, sd):
    return -np.inf if x <= 0 else -0.5 * np.square(x / sd)


def invgamma_from_bounds(l, u):
    """priors.py:30-32."""
    alpha = 2.0 + 9.0 * np.square((l + u) / (u - l))
    return alpha, 0.5 * (alpha - 1.0) * (l + u)


def prior_lpdf(spec, x):
    """spec = ('invgamma', alpha, beta) | ('halfnormal', sd)."""
    if spec[0] == "invgamma":
        return invgamma_lpdf(x, spec[1], spec[2])
    if spec[0] == "halfnormal":
        return halfnormal_lpdf(x, spec[1])
    raise ValueError(spec)


def prior_dlpdf(spec, x):
    """d lpdf / dx (closed form of priors.py:27, :50)."""
    if spec[0] == "invgamma":
        return -(spec[1] + 1.0) / x + spec[2] / (x * x)
    return -x / (spec[1] ** 2)


def unpack_tparams(model: Model, tparams, fix_R=False):
    """gpcsd1d.py:160-174 / gpcsd2d.py:185-199 -- log-space vector -> a new Model."""
    tparams = np.asarray(tparams, dtype=np.float64)
    ns = 1 if model.dim == 1 else 2
    R = model.R if fix_R else np.exp(tparams[0]) * 100.0
    ells = tuple(np.exp(tparams[1 + k]) * 100.0 for k in range(ns))
    p = 1 + ns
    temporal = []
    for kind, _, _ in model.temporal:
        temporal.append((kind, np.exp(tparams[p]), np.exp(tparams[p + 1])))
        p += 2
    sig2n = np.exp(tparams[p]) if (np.isscalar(model.sig2n) or np.ndim(model.sig2n) == 0) else np.exp(tparams[p:])
    return Model(model.dim, model.spatial, model.t, R, ells, temporal, sig2n, model.eps)


def pack_tparams(model: Model):
    v = [np.log(model.R / 100.0)] + [np.log(e / 100.0) for e in model.ells]
    for _, ell, s2 in model.temporal:
        v += [np.log(ell), np.log(s2)]
    v += list(np.atleast_1d(np.log(model.sig2n)))
    return np.array(v, dtype=np.float64)


def obj_fun(model: Model, lfp, tparams, priors, fix_R=False):
    """nll = -(loglik + sum lpdf).  ``priors`` is a flat list of specs in tparams order."""
    m = unpack_tparams(model, tparams, fix_R)
    vals = [m.R] + list(m.ells)
    for _, ell, s2 in m.temporal:
        vals += [ell, s2]
    vals += list(np.atleast_1d(m.sig2n))
    lp = sum(prior_lpdf(s, v) for s, v in zip(priors, vals))
    return -(loglik(m, lfp) + lp)


# --------------------------------------------------------------------------------------------
# closed-form gradient of loglik w.r.t. the natural parameters
# --------------------------------------------------------------------------------------------
def _dKs_contract(model: Model, G):
    """<G, dKs/dR>, <G, dKs/dell_k> for symmetric G (nx,nx); derivative of compKphi_{1d,2d}."""
    sp = model.spatial
    if model.dim == 1:
        ell = model.ells[0]
        g = sp.gl_x[None, :]
        d = (g - sp.x) / model.R
        A = sp.gl_w[None, :] * (np.sqrt(d * d + 1) - np.abs(d))
        dA = sp.gl_w[None, :] * (d * d / np.sqrt(d * d + 1) - np.abs(d)) * (-1.0 / model.R)
        dd = g.T - g
        Kg = np.exp(-0.5 * np.square(dd / ell))
        dKg = [Kg * np.square(dd) / ell ** 3]
    else:
        w = sp.delta_w(sp.x)
        Re = model.R + model.eps
        s = np.sqrt(Re ** 2 + w ** 2)
        A = sp.w_prod[None, :] * (np.log(Re + s) - np.log(model.eps + np.sqrt(model.eps ** 2 + w ** 2)))
        dA = sp.w_prod[None, :] * ((1.0 + Re / s) / (Re + s))
        sq1 = np.square(sp.grid1[:, None] - sp.grid1[None, :])
        sq2 = np.square(sp.grid2[:, None] - sp.grid2[None, :])
        Kg = np.exp(-0.5 * sq1 / model.ells[0] ** 2) * np.exp(-0.5 * sq2 / model.ells[1] ** 2)
        dKg = [Kg * sq1 / model.ells[0] ** 3, Kg * sq2 / model.ells[1] ** 3]
    GA = G @ A
    dR = 2.0 * np.sum(dA * (GA @ Kg))
    H = A.T @ GA
    return dR, [np.sum(H * dk) for dk in dKg]


def loglik_and_grad(model: Model, lfp, eigh=np.linalg.eigh):
    """loglik and d loglik / d(R, ells..., (ell_t, sigma2_t)..., sig2n[...]) in natural units.

    Derivation: with A_r = Qs^T Y_r Qt, D_ij = ls_i lt_j + s_i, B_r = A_r / D,
      Ms = sum_r B_r diag(lt) B_r^T,  Ns = sum_r B_r B_r^T,  Mt = sum_r B_r^T diag(ls) B_r,
      Dbar = -N/2 / D + 1/2 sum_r B_r^2,
      dL/dKs = Qs [ diag(sum_j Dbar_ij lt_j) + offdiag(1/2 Ms + 1/2 (s_i-s_i')/(ls_i-ls_i') Ns) ] Qs^T
      dL/dKt = Qt [ diag(sum_i Dbar_ij ls_i) + offdiag(1/2 Mt) ] Qt^T
      dL/ds_i = sum_j Dbar_ij.
    For scalar noise the (s_i - s_i') term vanishes.  This is the exact derivative of the reference's
    formula, including the eigen-index noise quirk (util:54-57)."""
    lfp = np.atleast_3d(lfp)
    nx, nt, N = lfp.shape
    Ks = model.Ks(jitter=True)
    Kt = model.Kt()
    Qs, Qt, Dv, ls, lt = comp_eig_D(Ks, Kt, model.sig2n, eigh)
    D = Dv.reshape(nx, nt)
    vec_noise = not (np.isscalar(model.sig2n) or np.ndim(model.sig2n) == 0)
    s = np.asarray(model.sig2n, dtype=np.float64) if vec_noise else np.full(nx, float(model.sig2n))
    A = (Qs.T @ lfp.reshape(nx, nt * N)).reshape(nx, nt, N)
    A = np.matmul(Qt.T[None, :, :], A)                                           # A_a = Qt^T Z_a
    B = A / D[:, :, None]
    ll = float(-0.5 * N * np.sum(np.log(D)) - 0.5 * np.sum(A * B))
    Bsq = np.sum(B * B, axis=2)
    Dbar = -0.5 * N / D + 0.5 * Bsq
    B2 = B.reshape(nx, nt * N)
    Ms = (B * lt[None, :, None]).reshape(nx, nt * N) @ B2.T                      # sum_{j,r} lt_j B_ajr B_bjr
    Bt = np.ascontiguousarray(B.transpose(1, 0, 2)).reshape(nt, nx * N)
    Mt = (Bt.reshape(nt, nx, N) * ls[None, :, None]).reshape(nt, nx * N) @ Bt.T  # sum_{a,r} ls_a B_ajr B_akr
    Xs = 0.5 * Ms
    if vec_noise:
        Ns = B2 @ B2.T
        dl = ls[:, None] - ls[None, :]
        ds = s[:, None] - s[None, :]
        with np.errstate(divide="ignore", invalid="ignore"):
            ratio = np.where(dl != 0, ds / dl, 0.0)
        Xs = Xs + 0.5 * ratio * Ns
    np.fill_diagonal(Xs, Dbar @ lt)
    Xt = 0.5 * Mt
    np.fill_diagonal(Xt, ls @ Dbar)
    Gs = Qs @ Xs @ Qs.T
    Gs = 0.5 * (Gs + Gs.T)
    Gt = Qt @ Xt @ Qt.T
    dR, dells = _dKs_contract(model, Gs)
    grad = [dR] + dells
    t = np.asarray(model.t, dtype=np.float64).reshape(-1, 1)
    dist = t - t.T
    for kind, ell, s2 in model.temporal:
        K = compute_Kt(kind, ell, s2, t)
        if kind == KIND_SE:
            dK = K * np.square(dist) / ell ** 3
        else:
            dK = K * np.abs(dist) / ell ** 2
        grad += [np.sum(Gt * dK), np.sum(Gt * K) / s2]
    dsv = np.sum(Dbar, axis=1)
    grad += list(dsv) if vec_noise else [np.sum(dsv)]
    return ll, np.array(grad, dtype=np.float64)


def obj_and_grad(model: Model, lfp, tparams, priors, fix_R=False):
    """nll and d nll / d tparams (log-space chain rule of gpcsd1d.py:160-174)."""
    m = unpack_tparams(model, tparams, fix_R)
    ll, g = loglik_and_grad(m, lfp)
    vals = [m.R] + list(m.ells)
    for _, ell, s2 in m.temporal:
        vals += [ell, s2]
    vals += list(np.atleast_1d(m.sig2n))
    vals = np.array(vals, dtype=np.float64)
    lp = sum(prior_lpdf(s, v) for s, v in zip(priors, vals))
    dlp = np.array([prior_dlpdf(s, v) for s, v in zip(priors, vals)])
    gt = -(g + dlp) * vals            # d value / d tparam = value for every exp-transform
    if fix_R:
        gt[0] = 0.0
    return -(ll + lp), gt


# --------------------------------------------------------------------------------------------
# prediction
# --------------------------------------------------------------------------------------------
def mykron(A, B):
    """utility_functions.py:35-42."""
    a1, a2 = A.shape
    b1, b2 = B.shape
    return (A[:, None, :, None] * B[None, :, None, :]).reshape(a1 * b1, a2 * b2)


def predict_dense(model: Model, lfp, z, tstar, kind="csd"):
    """gpcsd1d.py:248-293 / gpcsd2d.py:289-334 -- literal dense restatement (small shapes only).
    Returns dict with csd_pred, csd_pred_list, lfp_pred, lfp_pred_list as requested by ``kind``."""
    lfp = np.atleast_3d(lfp)
    nx, nt, N = lfp.shape
    nz, nts = z.shape[0], tstar.shape[0]
    yvec = np.reshape(lfp, (nx * nt, N))
    Qs, Qt, D, _, _ = comp_eig_D(model.Ks(jitter=False), model.Kt(), model.sig2n)   # no jitter: 1d:258
    ktmp = mykron(Qs, Qt)
    invy = np.dot(np.linalg.multi_dot([ktmp, np.diag(1.0 / D), ktmp.T]), yvec)
    out = {}
    cross = {}
    if kind in ("both", "csd"):
        cross["csd"] = model.Kphig(z)
    if kind in ("both", "lfp"):
        cross["lfp"] = model.Ks(xp=z)
    for name, Kc in cross.items():
        tot = np.zeros((nz, nts, N))
        parts = []
        for knd, ell, s2 in model.temporal:
            Kts = compute_Kt(knd, ell, s2, tstar, model.t)
            tmp = np.reshape(np.dot(mykron(Kc, Kts).T, invy), (nz, nts, N))
            parts.append(tmp)
            tot += tmp
        out[name + "_pred"] = tot
        out[name + "_pred_list"] = parts
    return out


def predict_kron(model: Model, lfp, z, tstar, kind="csd", eigh=np.linalg.eigh):
    """Same posterior mean in Kronecker form (never forms an (nx nt)^2 matrix):
    out_k[:, :, r] = (Kc^T Qs) ((Qs^T Y_r Qt) / D) (Qt^T Kt*_k)   with Kt*_k = Kt_k(t*, t) applied
    from the right exactly as ``mykron(Kc, Kt*).T @ invy`` does (needs len(t*) == len(t))."""
    lfp = np.atleast_3d(lfp)
    nx, nt, N = lfp.shape
    if tstar.shape[0] != nt:
        raise ValueError("shapes (%d,%d) and (%d,%d) not aligned" % (z.shape[0] * nt, nx * tstar.shape[0], nx * nt, N))
    Qs, Qt, D, _, _ = comp_eig_D(model.Ks(jitter=False), model.Kt(), model.sig2n, eigh)
    A = np.einsum("ia,ijr->ajr", Qs, lfp, optimize=True)
    A = np.einsum("ajr,jb->abr", A, Qt, optimize=True)
    B = A / D.reshape(nx, nt)[:, :, None]
    out = {}
    cross = {}
    if kind in ("both", "csd"):
        cross["csd"] = model.Kphig(z)
    if kind in ("both", "lfp"):
        cross["lfp"] = model.Ks(xp=z)
    for name, Kc in cross.items():
        V = np.einsum("za,abr->zbr", Kc.T @ Qs, B, optimize=True)
        parts = []
        for knd, ell, s2 in model.temporal:
            Pt = Qt.T @ compute_Kt(knd, ell, s2, tstar, model.t)
            parts.append(np.einsum("zbr,bj->zjr", V, Pt, optimize=True))
        out[name + "_pred"] = sum(parts[1:], parts[0].copy())
        out[name + "_pred_list"] = parts
    return out


# --------------------------------------------------------------------------------------------
# per-trial evoked-shift objective (auditory_lfp/fit_mean_function.py:304-321)
# --------------------------------------------------------------------------------------------
def shift_objective(lfp_trial, mu, t, tau, Qs, Qt, Dvec, mutau=0.0, sigtau=10.0):
    """Literal restatement of ``obj_fun(tau, lfp_trial)`` (fit_mean_function.py:311-321).
    lfp_trial (nx, nt); mu (nx, nt, nseg+1): background [:, :, 0] plus one evoked component per segment, each shifted in time
    by tau[i-1] through scipy interp1d(axis=1, fill_value="extrapolate") (:308); returns the scalar nll."""
    import scipy.interpolate
    ts = np.asarray(t, dtype=np.float64).squeeze()
    tau = np.asarray(tau, dtype=np.float64)
    nseg = mu.shape[2] - 1
    mu_new = np.copy(mu[:, :, 0])
    for i in range(1, nseg + 1):
        f = scipy.interpolate.interp1d(ts, mu[:, :, i], axis=1, fill_value="extrapolate")
        mu_new += f(ts + tau[i - 1])
    resid = lfp_trial - mu_new
    alpha = np.reshape(np.linalg.multi_dot([Qs.T, resid, Qt]), -1)
    quad = -0.5 * np.sum(alpha ** 2 / Dvec)
    nll = -1.0 * np.squeeze(quad)
    nll += -np.sum(-0.5 * np.square((tau - mutau) / sigtau))
    return float(nll)


def shift_objective_grad(lfp_trial, mu, t, tau, Qs, Qt, Dvec, mutau=0.0, sigtau=10.0):
    """Closed-form d nll / d tau of ``shift_objective`` (the reference lets scipy finite-difference it):
    d/dtau_s = - sum_ij (K^-1 resid)_ij * slope_s(i, t_j + tau_s) + (tau_s - mutau) / sigtau^2, K^-1 resid = Qs (alpha/D) Qt^T,
    slope = derivative of the piecewise-linear interpolant (end intervals extrapolate)."""
    ts = np.asarray(t, dtype=np.float64).squeeze()
    tau = np.asarray(tau, dtype=np.float64)
    nx, nt = lfp_trial.shape
    nseg = mu.shape[2] - 1
    mu_new = np.copy(mu[:, :, 0])
    slopes = []
    for i in range(1, nseg + 1):
        q = ts + tau[i - 1]
        k = np.clip(np.searchsorted(ts, q, side="right") - 1, 0, nt - 2)
        sl = (mu[:, k + 1, i] - mu[:, k, i]) / (ts[k + 1] - ts[k])[None, :]
        mu_new += mu[:, k, i] + sl * (q - ts[k])[None, :]
        slopes.append(sl)
    resid = lfp_trial - mu_new
    B = (Qs.T @ resid @ Qt) / np.reshape(Dvec, (nx, nt))
    V = Qs @ B @ Qt.T
    g = np.array([-np.sum(V * sl) for sl in slopes]) + (tau - mutau) / sigtau ** 2
    return g
