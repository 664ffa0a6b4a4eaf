"""Device-resident GPCSD engine: orchestrates the C-ABI kernels of libgpcsd_b200.so for one model.

Replaces the numpy/scipy/autograd arithmetic of the reference's ``loglik`` (gpcsd1d.py:113-128,
gpcsd2d.py:136-151), of ``grad(obj_fun)`` (gpcsd1d.py:211, gpcsd2d.py:250) and of ``predict``
(gpcsd1d.py:248-293, gpcsd2d.py:289-334).  PyTorch is used only for device memory, streams and
``torch.distributed``; every flop of the path runs in the hand-written sm_100a kernels (plus cuSOLVER
syevd for the two small eigendecompositions, as BASELINE.json's north_star allows).

Data layout in HBM (all FP64, row-major, trial index fastest exactly like the reference's C-order
``lfp[nx, nt, ntrials]``):
    Y, Z, Bm : [nx][nt][ldn]   ldn = ntrials_local rounded up to 8 (64-byte rows, zero padded)
    QsT, QtT : eigenvector matrices transposed (= cuSOLVER's column-major output), Qs, Qt row-major copies
    rD       : [nx][ld(nt)] reciprocal Kronecker eigenvalues 1/(ls_i lt_j + sig2n_i)

Trial sharding (torch.distributed): every rank holds a contiguous slab of trials; the small factors are
replicated (recomputed deterministically on every rank); partial (loglik, gradient) vectors are summed
with ONE all-reduce of P+1 doubles per evaluation.  ``predict`` needs no collective.
"""
import ctypes
from dataclasses import dataclass
from typing import List, Tuple, Union

import numpy as np
import torch

from . import _lib as L
from .parallel import TrialShard

F64 = torch.float64


def _even(n):
    return (int(n) + 1) // 2 * 2


def _ld8(n):
    return (int(n) + 7) // 8 * 8


@dataclass
class HyperParams:
    """Natural-unit hyperparameters read from a GPCSD1D/GPCSD2D object."""
    R: float
    ells: Tuple[float, ...]
    temporal: List[Tuple[int, float, float]]          # (kind, ell, sigma2) in temporal_cov_list order
    sig2n: Union[float, np.ndarray]                   # scalar or per spatial-eigen-index vector (util:54-57)
    eps: float = 0.0

    @property
    def vector_noise(self):
        return np.ndim(self.sig2n) > 0

    def n_params(self):
        return 1 + len(self.ells) + 2 * len(self.temporal) + (len(self.sig2n) if self.vector_noise else 1)


class KronEngine:
    """One model's geometry + LFP shard on one GPU."""

    def __init__(self, dim, x, t, quad, device=None, group=None, jitter=None):
        if not torch.cuda.is_available():
            raise L.GpcsdLibraryError("gpcsd_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        L.load()
        self.dim = int(dim)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.shard = TrialShard(group)
        self.jitter = (1e-8 if dim == 1 else 1e-7) if jitter is None else jitter   # gpcsd1d.py:17 / gpcsd2d.py:16
        self._ws = {}
        self._qcache = {}
        self.set_geometry(x, t, quad)
        self.Y = None
        self.ntrials_total = 0
        self.ntrials = 0
        self.n_launches = 0
        self._sides = None
        self._yf, self._yf_version, self._lfp_version = None, -1, 0
        self._copy_stream = None
        self._y_ready = None
        self.timers = None      # {abi_name: [(start_event, end_event), ...]} when bench.py profiles a kernel

    def share_data_with(self, other):
        """Evaluate on `other`'s uploaded LFP (no copy): lets several engines -- one per concurrent multi-start restart,
        each with its own workspaces and streams -- work on the same device-resident data."""
        if (other.nx, other.nt) != (self.nx, self.nt):
            raise ValueError("engines have different geometry")
        self.Y, self.ntrials, self.ntrials_total, self.ldn = other.Y, other.ntrials, other.ntrials_total, other.ldn
        self._y_ready = other._y_ready
        self._lfp_version += 1

    # ------------------------------------------------------------------ plumbing
    def _dev(self, arr):
        return torch.as_tensor(np.ascontiguousarray(arr, dtype=np.float64)).to(self.device)

    def _buf(self, name, *shape, dtype=F64):
        """Cached, zero-initialised device workspace."""
        key = (name, tuple(int(s) for s in shape), dtype)
        b = self._ws.get(key)
        if b is None:
            b = torch.zeros(tuple(int(s) for s in shape), dtype=dtype, device=self.device)
            self._ws[key] = b
        return b

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _q(self, name, *args):
        """Workspace-size queries are pure functions of their arguments: ask the library once per shape."""
        key = (name,) + args
        v = self._qcache.get(key)
        if v is None:
            v = self._qcache[key] = L.query(name, *args)
        return v

    # kernels of OURS launched per ABI call (cuSOLVER's own launches inside gpcsd_eigh are not counted)
    _LAUNCHES = {"gpcsd_project_quad": 2, "gpcsd_wsyrk": 2, "gpcsd_eig_D": 2, "gpcsd_kt_grad": 2, "gpcsd_dot": 2,
                 "gpcsd_eigh_dc": 3}

    def _call(self, name, *args, tag=None):
        self.n_launches += self._LAUNCHES.get(name, 1)
        timers = self.timers.get(tag or name) if self.timers else None
        if timers is None:
            return L.call(name, *args)
        # bench.py roofline: CUDA events on the launching stream around this ABI call
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = L.call(name, *args)
        e1.record()
        timers.append((e0, e1))
        return rc

    @staticmethod
    def _p(t, offset=0):
        return t.data_ptr() + 8 * int(offset)

    def gemm(self, transB, M, N, K, A, lda, sA, B, ldb, sB, C, ldc, sC, batch=1):
        self._call("gpcsd_dgemm", int(transB), int(M), int(N), int(K), self._p(A), lda, sA, self._p(B), ldb, sB,
                   self._p(C), ldc, sC, int(batch), self._stream())

    # ------------------------------------------------------------------ geometry / data
    def set_geometry(self, x, t, quad):
        """x: (nx,1) [1-D] or (nx,2) [2-D]; t: (nt,1); quad: GL nodes/weights on [a,b] (covariances.py:22-27,
        114-124).  Quadrature axes are padded to an even length with ZERO-weight nodes so every operand row
        is 16-byte aligned; zero weights contribute exactly 0 to every quadrature sum."""
        x = np.asarray(x, dtype=np.float64)
        self.nx = x.shape[0]
        self.x_host = x.reshape(-1) if self.dim == 1 else x.reshape(self.nx, 2)
        self._plans = {}
        self.t_host = np.asarray(t, dtype=np.float64).reshape(-1)
        self.nt = self.t_host.shape[0]
        self.t_dev = self._dev(self.t_host)
        # uniform to floating-point rounding (linspace / arange grids): then every stationary Kt is Toeplitz
        dt = np.diff(self.t_host)
        self.t_uniform = bool(self.nt >= 3 and np.all(dt > 0) and
                              np.max(np.abs(dt - dt[0])) <= 64 * np.finfo(np.float64).eps * np.max(np.abs(self.t_host)))
        if self.dim == 1:
            gx, gw = np.asarray(quad["gl_x"], dtype=np.float64), np.asarray(quad["gl_w"], dtype=np.float64)
            if len(gx) % 2:
                gx, gw = np.append(gx, gx[-1]), np.append(gw, 0.0)
            self.G = len(gx)
            self.quad_host = dict(gl_x=gx, gl_w=gw)
            self.gl_x, self.gl_w = self._dev(gx), self._dev(gw)
            self.x_dev = self._dev(x.reshape(-1))
        else:
            g1, w1 = np.asarray(quad["gl_x1"], dtype=np.float64), np.asarray(quad["gl_w1"], dtype=np.float64)
            g2, w2 = np.asarray(quad["gl_x2"], dtype=np.float64), np.asarray(quad["gl_w2"], dtype=np.float64)
            if len(g2) % 2:
                g2, w2 = np.append(g2, g2[-1]), np.append(w2, 0.0)
            self.G1, self.G2 = len(g1), len(g2)
            self.G = self.G1 * self.G2
            self.quad_host = dict(gl_x1=g1, gl_w1=w1, gl_x2=g2, gl_w2=w2)
            self.gl_x1, self.gl_w1, self.gl_x2, self.gl_w2 = self._dev(g1), self._dev(w1), self._dev(g2), self._dev(w2)
            self.x_dev = self._dev(x.reshape(self.nx, 2))
        self.ldx = _even(self.nx)
        self.ldt = _even(self.nt)
        self.s_pairs = self._find_reflection_pairs(x, quad)

    def _find_reflection_pairs(self, x, quad):
        """Fixed-point-free involution pi of the electrode sites induced by the point reflection s -> c - s about the
        centre of the integration box (c = a + b per dimension; the Gauss-Legendre nodes are symmetric about it).  If
        the site set is invariant, Ks[pi(i)][pi(j)] == Ks[i][j] and the spatial eigenproblem splits into two
        independent ones of half the order.  Returns (ra, rb) device int32 arrays or None."""
        nx = self.nx
        self.s_pairs_host = None
        if nx < 4 or nx % 2:
            return None
        if self.dim == 1:
            g = np.asarray(quad["gl_x"], dtype=np.float64)
            pts, c = x.reshape(nx, 1), np.array([g.min() + g.max()])
        else:
            g1, g2 = np.asarray(quad["gl_x1"], dtype=np.float64), np.asarray(quad["gl_x2"], dtype=np.float64)
            pts, c = x.reshape(nx, 2), np.array([g1.min() + g1.max(), g2.min() + g2.max()])
        scale = max(float(np.max(np.abs(pts))), float(np.max(np.abs(c))), 1e-300)
        key = lambda p: tuple(np.round(p / scale, 11))
        index = {}
        for i in range(nx):
            k = key(pts[i])
            if k in index:
                return None            # duplicate sites
            index[k] = i
        pi = np.empty(nx, dtype=np.int64)
        for i in range(nx):
            j = index.get(key(c - pts[i]))
            if j is None or j == i:
                return None
            pi[i] = j
        if not np.array_equal(pi[pi], np.arange(nx)):
            return None
        # exactness guard: the reflected coordinates must agree to rounding, not just to the matching tolerance
        if np.max(np.abs((c - pts) - pts[pi])) > 64 * np.finfo(np.float64).eps * scale:
            return None
        ra = np.array([i for i in range(nx) if i < pi[i]], dtype=np.int32)
        rb = pi[ra].astype(np.int32)
        self.s_pairs_host = (ra, rb)
        return (torch.from_numpy(ra).to(self.device), torch.from_numpy(rb).to(self.device))

    def set_lfp(self, lfp, local=False):
        """Upload this rank's slab of trials.  lfp: host (nx, nt, ntrials) float64 (C order; numpy array or
        torch tensor, pinned memory makes the copy an async DMA) or a CUDA tensor of that shape.
        local=False: ``lfp`` holds ALL trials and this rank takes its contiguous slab;
        local=True : ``lfp`` already is this rank's slab (the global trial count is all-reduced).
        Returns the bytes copied host->device."""
        if isinstance(lfp, torch.Tensor):
            src = lfp if lfp.dim() == 3 else lfp.reshape(lfp.shape[0], lfp.shape[1], -1)
            if src.dtype != F64:
                raise TypeError("lfp tensor must be float64")
        else:
            src = torch.from_numpy(np.ascontiguousarray(np.atleast_3d(np.asarray(lfp, dtype=np.float64))))
        ntot = src.shape[2]
        if src.shape[0] != self.nx or src.shape[1] != self.nt:
            raise ValueError("lfp shape %s does not match (nx=%d, nt=%d)" % (tuple(src.shape), self.nx, self.nt))
        if local:
            lo, hi = 0, ntot
            ntot = int(round(self.shard.allreduce_sum(np.array([float(ntot)]), self.device)[0]))
        else:
            lo, hi = self.shard.bounds(ntot)
        n = hi - lo
        ldn = _ld8(max(n, 1))
        if self.Y is None or self.Y.shape != (self.nx, self.nt, ldn):
            self.Y = torch.zeros((self.nx, self.nt, ldn), dtype=F64, device=self.device)
            self._ws.clear()                                   # workspaces sized by the old trial count are dropped
        slab = src[:, :, lo:hi]
        if src.is_cuda:
            self.Y[:, :, :n].copy_(slab)
            self._y_ready = None
            nbytes = 0
        else:
            # Upload on a dedicated copy stream: the covariance build and the eigendecompositions that open
            # every evaluation do not touch Y, so the DMA (pinned host memory: ~55 GB/s) hides behind them; the
            # projection GEMM waits on `_y_ready`.
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream(device=self.device)
            cur = torch.cuda.current_stream(self.device)
            self._copy_stream.wait_stream(cur)                 # earlier readers of Y are ordered before the overwrite
            with torch.cuda.stream(self._copy_stream):
                dst = self.Y if ldn == n else self.Y[:, :, :n]
                dst.copy_(slab if slab.is_contiguous() else slab.contiguous(), non_blocking=True)
                self._y_ready = torch.cuda.Event()
                self._y_ready.record(self._copy_stream)
            nbytes = slab.numel() * 8
        self.ntrials_total, self.ntrials, self.ldn = ntot, n, ldn
        self._lfp_version += 1
        return nbytes

    # ------------------------------------------------------------------ covariance construction
    def _fwd_weights(self, pts_dev, npts, hp, want_dA, tag):
        A = self._buf("A_" + tag, npts, self.G)
        dA = self._buf("dA_" + tag, npts, self.G) if want_dA else None
        if self.dim == 1:
            self._call("gpcsd_fwd_weights_1d", npts, self._p(pts_dev), self.G, self._p(self.gl_x), self._p(self.gl_w),
                       float(hp.R), self._p(A), self._p(dA) if want_dA else None, self.G, self._stream())
        else:
            self._call("gpcsd_fwd_weights_2d", npts, self._p(pts_dev), self.G1, self.G2, self._p(self.gl_x1),
                       self._p(self.gl_w1), self._p(self.gl_x2), self._p(self.gl_w2), float(hp.R), float(hp.eps),
                       self._p(A), self._p(dA) if want_dA else None, self.G, self._stream())
        return A, dA

    def _se(self, name, a, b, ell, deriv=0):
        na, nb = a.shape[0], b.shape[0]
        out = self._buf(name, na, _even(nb))
        self._call("gpcsd_se_matrix", na, self._p(a), nb, self._p(b), float(ell), 1.0, int(deriv), self._p(out),
                   _even(nb), self._stream())
        return out

    def _quad_kernels(self, hp, deriv_axis=None, tag=""):
        """CSD SE kernel on the quadrature grid: 1-D -> [Kg]; 2-D -> Kronecker factors [K1, K2]
        (covariances.py:89, 216: exp(-sq1/2l1^2) * exp(-sq2/2l2^2) on the x1-major product grid)."""
        if self.dim == 1:
            return [self._se("Kg" + tag, self.gl_x, self.gl_x, hp.ells[0], deriv=int(deriv_axis == 0))]
        return [self._se("Kg1" + tag, self.gl_x1, self.gl_x1, hp.ells[0], deriv=int(deriv_axis == 0)),
                self._se("Kg2" + tag, self.gl_x2, self.gl_x2, hp.ells[1], deriv=int(deriv_axis == 1))]

    def _apply_quad_kernel(self, X, nrows, kern, out_name):
        """out = X (nrows x G) * Kg, Kg given by _quad_kernels (dense 1-D, Kronecker 2-D)."""
        out = self._buf(out_name, nrows, self.G)
        if self.dim == 1:
            Kg = kern[0]
            self.gemm(0, nrows, self.G, self.G, X, self.G, 0, Kg, Kg.shape[1], 0, out, self.G, 0)
            return out
        K1, K2 = kern
        G1, G2 = self.G1, self.G2
        T = self._buf("kron_tmp", nrows, self.G)
        # T[(i,a), b'] = sum_b X[(i,a), b] K2[b, b']
        self.gemm(0, nrows * G1, G2, G2, X, G2, 0, K2, K2.shape[1], 0, T, G2, 0)
        # out_i[a', b'] = sum_a K1[a', a] T_i[a, b']   (K1 symmetric), batched over rows i
        self.gemm(0, G1, G2, G1, K1, K1.shape[1], 0, T, G2, self.G, out, G2, self.G, batch=nrows)
        return out

    def _spatial_cov(self, hp, jitter, want_grad):
        """Ks = (A Kg) A^T (+ jitter I)   covariances.py:74-96 / 204-232."""
        A, dA = self._fwd_weights(self.x_dev, self.nx, hp, want_grad, "x")
        kern = self._quad_kernels(hp)
        U = self._apply_quad_kernel(A, self.nx, kern, "U")
        Ks = self._buf("Ks", self.nx, self.ldx)
        self.gemm(1, self.nx, self.nx, self.G, U, self.G, 0, A, self.G, 0, Ks, self.ldx, 0)
        if jitter:
            self._call("gpcsd_add_diag", self.nx, self._p(Ks), self.ldx, float(self.jitter), self._stream())
        return Ks, A, dA, U, kern

    def _temporal_spec(self, temporal):
        return (len(temporal), L.c_int_array([k for k, _, _ in temporal]), L.c_double_array([e for _, e, _ in temporal]),
                L.c_double_array([s for _, _, s in temporal]))

    def _temporal_cov(self, hp):
        Kt = self._buf("Kt", self.nt, self.ldt)
        ntc, kinds, ells, s2 = self._temporal_spec(hp.temporal)
        self._call("gpcsd_kt_build", self.nt, self._p(self.t_dev), self.nt, self._p(self.t_dev), ntc, kinds, ells, s2,
                   self._p(Kt), self.ldt, self._stream())
        return Kt

    DC_EIGH_MAX = 256           # order limit of the in-house cluster eigensolver (gpcsd_eigh_dc, csrc/gpcsd_eig.cu)

    def _eigh(self, K, n, ld, tag, QT=None, W=None):
        """Eigen-factors of one symmetric matrix (eigenvectors as rows, eigenvalues ascending): the in-house cluster
        solver up to order 256, cuSOLVER syevd above."""
        if 3 <= n <= self.DC_EIGH_MAX:
            QTs, Ws, info = self._eigh_dc(K, n, ld, 1, tag, QT=QT, W=W)
            return QTs[0], Ws[0], info
        QT = self._buf("QT_" + tag, n, ld) if QT is None else QT
        W = self._buf("W_" + tag, n) if W is None else W
        nws = self._q("gpcsd_eigh_ws_doubles", n, ld)
        ws = self._buf("eigws_" + tag, max(nws, 1))
        info = self._buf("info_" + tag, 1, dtype=torch.int32)
        self._call("gpcsd_eigh", n, self._p(K), ld, self._p(QT), ld, self._p(W), self._p(ws), nws, info.data_ptr(),
                   self._stream())
        return QT, W, info

    def _eigh_dc(self, stack, n, ld, nmat, tag, QT=None, W=None):
        """`nmat` stacked symmetric matrices [nmat][n][ld] -> (QT [nmat][n][ld] rows = eigenvectors, W [nmat][n] ascending)
        in one launch sequence of the cluster solver (one 8-CTA cluster per matrix); the input is left untouched."""
        QT = self._buf("QTdc_" + tag, nmat, n, ld) if QT is None else QT.view(nmat, n, ld)
        W = self._buf("Wdc_" + tag, nmat, n) if W is None else W.view(nmat, n)
        nws = self._q("gpcsd_eigh_dc_ws_doubles", n, ld, nmat)
        ws = self._buf("eigdcws_" + tag, nws)
        info = self._buf("infodc_" + tag, nmat, dtype=torch.int32)
        self._call("gpcsd_eigh_dc", n, nmat, self._p(stack), ld, self._p(QT), ld, self._p(W), self._p(ws), nws, info.data_ptr(),
                   self._stream())
        if n >= 97:
            self.n_launches += 3      # H^T formation (2 kernels) + one more GEMM on the large-order path
        return QT, W, info

    def _eigh_pair(self, S, A, m, ld, tag, side):
        """Eigen-factors of two independent symmetric matrices of order m (leading dimension ld).  Order <= 256: one
        batched call of the cluster solver (S and A must then be the two slabs of one [2][m][ld] stack); larger: two
        syevd on two streams."""
        if 3 <= m <= self.DC_EIGH_MAX and A.data_ptr() == S.data_ptr() + 8 * m * ld:
            QT, W, info = self._eigh_dc(S, m, ld, 2, tag)
            return QT[0], W[0], QT[1], W[1], [info]
        main = torch.cuda.current_stream(self.device)
        ready = torch.cuda.Event()
        ready.record(main)
        with torch.cuda.stream(side):
            side.wait_event(ready)
            UaT, Wa, info_a = self._eigh(A, m, ld, tag + "a")
            done = torch.cuda.Event()
            done.record(side)
        UsT, Ws, info_s = self._eigh(S, m, ld, tag + "s")
        main.wait_event(done)
        return UsT, Ws, UaT, Wa, [info_s, info_a]

    def _eigh_spatial(self, Ks, allow_split):
        """Factor of order nx.  If the geometry has the point-reflection symmetry (and the eigenvalue order is not needed:
        scalar noise) the problem is split into two of order nx/2; orders <= 128 go through the batched small-matrix path
        (batch 2 -- batch 1 would fall back to the latency-bound syevd); everything else is one syevd."""
        nx, ld = self.nx, self.ldx
        self._s_blocks = None
        if allow_split and self.s_pairs is not None and nx >= 64:
            ra, rb = self.s_pairs
            m = nx // 2
            ldm = _even(m)
            stack = self._buf("ps_stack", 2, m, ldm)
            self._call("gpcsd_pairsym_split", nx, self._p(Ks), ld, ra.data_ptr(), rb.data_ptr(), self._p(stack), ldm,
                       self._p(stack, m * ldm), ldm, self._stream())
            UsT, Ws, UaT, Wa, infos = self._eigh_pair(stack[0], stack[1], m, ldm, "sp", self._side_streams()[2])
            QT, W = self._buf("QT_s", nx, ld), self._buf("W_s", nx)
            self._call("gpcsd_pairsym_assemble", nx, ra.data_ptr(), rb.data_ptr(), self._p(UsT), ldm, self._p(Ws),
                       self._p(UaT), ldm, self._p(Wa), self._p(QT), ld, self._p(W), self._stream())
            self._s_blocks = (UsT, UaT, ldm)
            return QT, W, infos
        QT, W, info = self._eigh(Ks, nx, ld, "s")
        return QT, W, [info]

    def _side_streams(self):
        if self._sides is None:
            self._sides = [torch.cuda.Stream(device=self.device) for _ in range(3)]
        return self._sides

    def _folded_lfp(self):
        """The uploaded LFP in the channel-folded basis of gpcsd_pairsym_fold; recomputed only after a new upload.  Runs on
        the caller's (side) stream, after the upload event."""
        if self._yf is None or self._yf_version != self._lfp_version or self._yf.shape != self.Y.shape:
            ra, rb = self.s_pairs
            if self._yf is None or self._yf.shape != self.Y.shape:
                self._yf = torch.zeros_like(self.Y)
            self._call("gpcsd_pairsym_fold", self.nx, ra.data_ptr(), rb.data_ptr(), self.nt * self.ldn, self._p(self.Y),
                       self._p(self._yf), self._stream())
            self._yf_version = self._lfp_version
        return self._yf

    def _t_split(self):
        """Uniform time grid: Kt is centrosymmetric, its eigenproblem splits into a symmetric and a skew half and -- in the
        folded time basis (gpcsd_centro_fold) -- Qt is block diagonal."""
        return bool(self.t_uniform and self.nt >= 32)

    FOLD_MIN_NT = 128           # below this the evaluation is launch-bound and the extra launches of the fold cost more

    def _t_fold(self):
        """Whether the trial data are moved to the folded time basis: the split applies and the temporal contractions are
        big enough for halving their flops to matter (configs[0]-sized problems are bound by the number of launches)."""
        return self._t_split() and self.nt >= self.FOLD_MIN_NT

    def _eigh_temporal(self, Kt):
        """Eigen-factors of Kt.  On a uniform time grid Kt is symmetric Toeplitz, hence centrosymmetric, and the
        order-nt problem splits exactly into two independent problems of order ~nt/2 (one batched call of the cluster
        solver up to nt = 512, two concurrent syevd above); otherwise one solve of order nt.  The eigenvalues come back unsorted in the split case (nothing downstream needs an order)."""
        nt, ldt = self.nt, self.ldt
        self._t_blocks = None
        if not self._t_split():
            QT, W, info = self._eigh(Kt, nt, ldt, "t")
            return QT, W, [info]
        m, ms = nt // 2, nt // 2 + (nt & 1)
        lds, lda = _even(ms), _even(m)
        if nt % 2 == 0 and 3 <= m <= self.DC_EIGH_MAX:
            # both halves have order m <= 256: one batched call of the cluster solver (two clusters side by side)
            stack = self._buf("cs_stack", 2, m, lds)
            self._call("gpcsd_centro_split", nt, self._p(Kt), ldt, self._p(stack), lds, self._p(stack, m * lds), lds,
                       self._stream())
            U, Wb, info = self._eigh_dc(stack, m, lds, 2, "t")
            QT, W = self._buf("QT_t", nt, ldt), self._buf("W_t", nt)
            self._call("gpcsd_centro_assemble", nt, self._p(U), lds, self._p(Wb), self._p(U, m * lds), lds,
                       self._p(Wb, m), self._p(QT), ldt, self._p(W), self._stream())
            self._t_blocks = (U[0], lds, U[1], lds)
            return QT, W, [info]
        S, A = self._buf("cs_S", ms, lds), self._buf("cs_A", m, lda)
        self._call("gpcsd_centro_split", nt, self._p(Kt), ldt, self._p(S), lds, self._p(A), lda, self._stream())
        main = torch.cuda.current_stream(self.device)
        side = self._side_streams()[1]
        ready = torch.cuda.Event()
        ready.record(main)
        with torch.cuda.stream(side):
            side.wait_event(ready)
            UaT, Wa, info_a = self._eigh(A, m, lda, "ta")
            done = torch.cuda.Event()
            done.record(side)
        UsT, Ws, info_s = self._eigh(S, ms, lds, "ts")
        main.wait_event(done)
        QT, W = self._buf("QT_t", nt, ldt), self._buf("W_t", nt)
        self._call("gpcsd_centro_assemble", nt, self._p(UsT), lds, self._p(Ws), self._p(UaT), lda, self._p(Wa),
                   self._p(QT), ldt, self._p(W), self._stream())
        self._t_blocks = (UsT, lds, UaT, lda)
        return QT, W, [info_s, info_a]

    def _given_factors(self, st, factors, Ydata=None):
        """Caller-supplied eigen-factors (Qs, ls, Qt, lt) -- host arrays with eigenvectors as COLUMNS, as np.linalg.eigh
        returns them in comp_eig_D (utility_functions.py:58-59) -- uploaded in place of the eigensolvers' output.  This is the
        kernel-level entry of SURVEY.md section 6: with identical factors on both sides everything downstream of comp_eig_D
        (projection, quadratic form, SYRKs, gradient cores, rotations, covariance-derivative contractions) must agree with
        the oracle to rounding, per-electrode noise included."""
        Qs, ls, Qt, lt = (np.asarray(a, dtype=np.float64) for a in factors)
        if Qs.shape != (self.nx, self.nx) or Qt.shape != (self.nt, self.nt) or ls.shape != (self.nx,) or lt.shape != (self.nt,):
            raise ValueError("factors must be (Qs (nx,nx), ls (nx,), Qt (nt,nt), lt (nt,))")
        self._s_blocks = self._t_blocks = None
        QsT, QtT = self._buf("QT_s", self.nx, self.ldx), self._buf("QT_t", self.nt, self.ldt)
        QsT[:, :self.nx].copy_(self._dev(Qs.T))
        QtT[:, :self.nt].copy_(self._dev(Qt.T))
        st["QsT"], st["ls"], st["QtT"], st["lt"] = QsT, self._dev(ls), QtT, self._dev(lt)
        st["infos"] = [torch.zeros(1, dtype=torch.int32, device=self.device)]
        Ydata = self.Y if Ydata is None else Ydata
        if Ydata is not None:
            if self._y_ready is not None:
                torch.cuda.current_stream(self.device).wait_event(self._y_ready)
            st["Z"] = self._buf("Z", self.nx, self.nt, self.ldn)
            row = self.nt * self.ldn
            self.gemm(0, self.nx, row, self.nx, QsT, self.ldx, 0, Ydata, row, 0, st["Z"], row, 0)

    def _factorize(self, hp, jitter, want_grad, factors=None, data=None):
        """Covariances -> eigen-factors -> 1/D and its reductions (comp_eig_D, utility_functions.py:44-64).
        The spatial and the temporal eigenproblems are independent and run on separate streams."""
        st = {}
        st["Ks"], st["A"], st["dA"], st["U"], st["kern"] = self._spatial_cov(hp, jitter, want_grad)
        Ydata = self.Y if data is None else data          # `data`: another [nx][nt][ldn] array to project (shift residuals)
        if factors is not None:
            self._given_factors(st, factors, Ydata)
            return self._finish_factorize(st, hp)
        main = torch.cuda.current_stream(self.device)
        side = self._side_streams()[0]
        ks_ready = torch.cuda.Event()
        ks_ready.record(main)
        with torch.cuda.stream(side):
            side.wait_event(ks_ready)
            st["QsT"], st["ls"], infos_s = self._eigh_spatial(st["Ks"], allow_split=not hp.vector_noise)
            if Ydata is not None:
                # Z = Qs^T Y needs the spatial factor only: run it here, underneath the (longer) temporal eigensolve
                if self._y_ready is not None:
                    side.wait_event(self._y_ready)
                if data is not None:
                    side.wait_stream(main)                   # `data` was produced on the main stream
                st["Z"] = self._buf("Z", self.nx, self.nt, self.ldn)
                row = self.nt * self.ldn
                if self._s_blocks is not None and self.ntrials > 0 and data is None:
                    # reflection-symmetric geometry: Qs is block diagonal on the channel-folded LFP (folded once per upload)
                    UsT, UaT, ldm = self._s_blocks
                    m = self.nx // 2
                    Yf = self._folded_lfp()
                    self._call("gpcsd_dgemm", 0, m, row, m, self._p(UsT), ldm, 0, self._p(Yf), row, 0, self._p(st["Z"]), row, 0,
                               1, self._stream())
                    self._call("gpcsd_dgemm", 0, m, row, m, self._p(UaT), ldm, 0, self._p(Yf, m * row), row, 0,
                               self._p(st["Z"], m * row), row, 0, 1, self._stream())
                else:
                    self.gemm(0, self.nx, row, self.nx, st["QsT"], self.ldx, 0, Ydata, row, 0, st["Z"], row, 0)
                if self._t_fold() and self.ntrials > 0:
                    # folded time basis: the temporal projection and the temporal SYRK then run on two blocks of order
                    # ~nt/2 (half the flops); also hidden underneath the temporal eigensolve
                    Zf = self._buf("Zf", self.nx, self.nt, self.ldn)
                    self._call("gpcsd_centro_fold", self.nx, self.nt, self.ldn, self._p(st["Z"]), self._p(Zf), self._stream())
                    st["Z"] = Zf
            s_done = torch.cuda.Event()
            s_done.record(side)
        st["Kt"] = self._temporal_cov(hp)
        st["QtT"], st["lt"], infos_t = self._eigh_temporal(st["Kt"])
        main.wait_event(s_done)
        st["infos"] = infos_s + infos_t
        return self._finish_factorize(st, hp)

    def _finish_factorize(self, st, hp):
        s_host = np.atleast_1d(np.asarray(hp.sig2n, dtype=np.float64))
        if len(s_host) not in (1, self.nx):
            raise ValueError("sig2n must be a scalar or have one entry per electrode")
        st["s"] = self._dev(s_host)
        st["rD"] = self._buf("rD", self.nx, self.ldt)
        st["sums"] = self._buf("res", 64)            # [0:2]=quad,bsq  [2:4]=sum log D, sum 1/D  [4:]=gradient dots
        st["rowA"], st["rowC"], st["rowL"] = self._buf("rowA", self.nx), self._buf("rowC", self.nx), self._buf("rowL", self.nx)
        st["colB"] = self._buf("colB", self.nt)
        self._call("gpcsd_eig_D", self.nx, self.nt, self._p(st["ls"]), self._p(st["lt"]), self._p(st["s"]), len(s_host),
                   self._p(st["rD"]), self.ldt, self._p(st["sums"], 2), self._p(st["rowA"]), self._p(st["rowC"]),
                   self._p(st["rowL"]), self._p(st["colB"]), self._stream())
        return st

    def _project(self, st):
        """Z = Qs^T Y (all trials), then per spatial eigen-index A_i = Qt^T Z_i with the fused /D + quadratic
        form epilogue (the hot loop gpcsd1d.py:124-126)."""
        if self.Y is None:
            raise RuntimeError("no LFP uploaded: call set_lfp first")
        nx, nt, ldn = self.nx, self.nt, self.ldn
        Z = st["Z"]                                  # computed on the spatial side stream by _factorize
        Bm = self._buf("Bm", nx, nt, ldn)
        # (the block calls of the folded basis may take another kernel path than order nt: size for all three orders)
        nws = max(self._q("gpcsd_project_quad_ws_doubles", nx, o, max(self.ntrials, 1)) for o in {nt, nt // 2, nt - nt // 2})
        part = self._buf("quad_part", nws)
        st["sums_b"] = None
        if self.ntrials > 0 and self._t_blocks is not None and self._t_fold():
            # Z is in the folded time basis: Qt^T Z_i = [Us^T Zf_i[:ms]; Ua^T Zf_i[ms:]]
            UsT, lds, UaT, lda = self._t_blocks
            m = nt // 2
            ms = nt - m
            st["sums_b"] = self._buf("res_b", 2)
            self._call("gpcsd_project_quad_strided", nx, ms, self.ntrials, self._p(UsT), lds, self._p(Z), ldn, nt * ldn,
                       self._p(st["rD"]), self.ldt, self._p(Bm), self._p(part), self._p(st["sums"], 0), self._stream())
            self._call("gpcsd_project_quad_strided", nx, m, self.ntrials, self._p(UaT), lda, self._p(Z, ms * ldn), ldn,
                       nt * ldn, self._p(st["rD"], ms), self.ldt, self._p(Bm, ms * ldn), self._p(part),
                       self._p(st["sums_b"], 0), self._stream())
        elif self.ntrials > 0:
            self._call("gpcsd_project_quad", nx, nt, self.ntrials, self._p(st["QtT"]), self.ldt, self._p(Z), ldn,
                       self._p(st["rD"]), self.ldt, self._p(Bm), self._p(part), self._p(st["sums"], 0), self._stream())
        else:
            st["sums"][0:2].zero_()
        st["Bm"] = Bm
        return Bm

    # ------------------------------------------------------------------ public evaluations
    def _check_info(self, st):
        if any(bool(torch.any(i != 0).item()) for i in st["infos"]):
            raise np.linalg.LinAlgError("Eigenvalues did not converge")

    def _theta_check_piece(self, hp):
        """Trial-sharded evaluations are only meaningful if every rank evaluates the SAME hyperparameters (ranks that seeded
        numpy differently would silently sum likelihood terms of different models).  Two extra entries ride on the
        evaluation's one all-reduce: c/world and c^2/world for a fixed weighted checksum c of the hyperparameters; after the
        sum they are mean(c) and mean(c^2), and a non-zero variance means the ranks disagree.  None when not sharded."""
        if not (self.shard.enabled and self.shard.world > 1):
            return None
        v = np.concatenate([[hp.R], np.asarray(hp.ells, dtype=np.float64), np.array([[e, s] for _, e, s in hp.temporal]).reshape(-1),
                            np.atleast_1d(np.asarray(hp.sig2n, dtype=np.float64))])
        c = float(np.dot(np.sin(1.0 + np.arange(v.size)), np.log(np.maximum(v, 1e-300))))
        return self._dev(np.array([c, c * c]) / self.shard.world)

    @staticmethod
    def _theta_check(flat, have):
        """Strip the checksum pair appended by _theta_check_piece and raise if the ranks' hyperparameters differ."""
        if not have:
            return flat
        mean, meansq = flat[-2], flat[-1]
        if abs(meansq - mean * mean) > 1e-9 * max(abs(meansq), 1e-300):
            raise RuntimeError("trial-sharded evaluation: the ranks hold different hyperparameters (checksum variance %.3e); "
                               "seed numpy identically on every rank or broadcast the parameters" % (meansq - mean * mean))
        return flat[:-2]

    @staticmethod
    def _info_piece(st):
        """The eigensolvers' info flags as one float64 device vector (>= 0 entries), so they travel with the result
        instead of costing a device->host read each."""
        return torch.cat([i.reshape(-1) for i in st["infos"]]).abs().to(F64)

    # ------------------------------------------------------------------ native plan: one ABI call per evaluation
    use_plan = True             # False: orchestrate the evaluation call by call from Python (the stepwise path below)

    @staticmethod
    def theta_of(hp):
        """Natural-unit hyperparameter vector in the plan's (and the gradient's) order."""
        return np.concatenate([[hp.R], np.asarray(hp.ells, dtype=np.float64),
                               np.array([[e, s] for _, e, s in hp.temporal], dtype=np.float64).reshape(-1),
                               np.atleast_1d(np.asarray(hp.sig2n, dtype=np.float64))])

    def _plan_for(self, hp, R=1):
        from .plan import EvalPlan
        nsig = self.nx if hp.vector_noise else 1
        if hp.vector_noise and len(hp.sig2n) != self.nx:
            raise ValueError("sig2n must be a scalar or have one entry per electrode")
        key = (tuple(k for k, _, _ in hp.temporal), nsig, float(hp.eps))
        pl = self._plans.get(key)
        if pl is None or pl.rmax < R:
            self._plans[key] = None
            pl = self._plans[key] = EvalPlan(self, key[0], nsig, eps=key[2], max_restarts=max(R, 1))
        return pl

    def max_batch(self, hp, want):
        """Largest restart batch (<= want) whose plan workspace fits in half of the free device memory."""
        free, _ = torch.cuda.mem_get_info(self.device)
        slab = 8 * self.nx * self.nt * self.ldn
        per = 2 * slab + 8 * (7 * self.nx * self.G + 8 * self.nx * self.ldx + 8 * self.nt * self.ldt) + (64 << 20)
        return int(max(1, min(want, (free // 2) // max(per, 1))))

    def loglik_grad_batch(self, hps, want_grad=True):
        """Evaluate R hyperparameter sets in ONE native call (restart-batched kernels, CUDA-graph replay).
        Returns (loglik (R,), gradient (R, P) in natural units, solver flags (R,): non-zero where numpy's eigh would raise)."""
        hps = list(hps)
        R = len(hps)
        # (cudaMemGetInfo costs up to milliseconds: only ask when a bigger plan workspace would have to be allocated)
        have = self._plans.get((tuple(k for k, _, _ in hps[0].temporal), self.nx if hps[0].vector_noise else 1, float(hps[0].eps)))
        chunk = R if (R == 1 or (have is not None and have.rmax >= R)) else self.max_batch(hps[0], R)
        outs = []
        for lo in range(0, R, chunk):
            part = hps[lo: lo + chunk]
            # batch sizes are rounded up to a power of two (the last vector repeated) so that a lock-step optimiser whose
            # active set shrinks restart by restart replays a handful of recorded graphs instead of capturing one per size
            n = len(part)
            npad = min(1 << (n - 1).bit_length(), max(chunk, n)) if n > 1 else 1
            pl = self._plan_for(part[0], npad)
            th = np.stack([self.theta_of(h) for h in part] + [self.theta_of(part[-1])] * (npad - n))
            outs.append(pl.evaluate(th, want_grad)[:n])
            self.n_launches += max(pl.last_launches(), 0)
        out = np.concatenate(outs, axis=0)
        P = out.shape[1] - 4
        return out[:, 0].copy(), out[:, 1:1 + P].copy(), out[:, 1 + P].copy()

    def loglik_grad_thetas(self, thetas, template, want_grad=True):
        """loglik_grad_batch for hyperparameters already laid out as rows of natural-unit theta vectors (the plan's order);
        `template` is any HyperParams of the same structure (temporal kernel kinds, scalar / per-electrode noise, eps)."""
        thetas = np.ascontiguousarray(thetas, dtype=np.float64)
        R = thetas.shape[0]
        key = (tuple(k for k, _, _ in template.temporal), self.nx if template.vector_noise else 1, float(template.eps))
        have = self._plans.get(key)
        chunk = R if (R == 1 or (have is not None and have.rmax >= R)) else self.max_batch(template, R)
        outs = []
        for lo in range(0, R, chunk):
            part = thetas[lo: lo + chunk]
            n = part.shape[0]
            npad = min(1 << (n - 1).bit_length(), max(chunk, n)) if n > 1 else 1
            pl = self._plan_for(template, npad)
            th = part if npad == n else np.concatenate([part, np.repeat(part[-1:], npad - n, axis=0)], axis=0)
            outs.append(pl.evaluate(th, want_grad)[:n])
            self.n_launches += max(pl.last_launches(), 0)
        out = np.concatenate(outs, axis=0)
        P = out.shape[1] - 4
        return out[:, 0].copy(), out[:, 1:1 + P].copy(), out[:, 1 + P].copy()

    def loglik(self, hp, factors=None):
        """Marginal log-likelihood (gpcsd1d.py:113-128 / gpcsd2d.py:136-151); all-reduced over trial shards.
        factors: optional caller-supplied (Qs, ls, Qt, lt), see _given_factors."""
        if self.use_plan and self.timers is None:
            if factors is not None:
                return float(self._plan_factors(hp, factors, False)[0])
            ll, _, flag = self.loglik_grad_batch([hp], want_grad=False)
            if flag[0] != 0:
                raise np.linalg.LinAlgError("Eigenvalues did not converge")
            return float(ll[0])
        st = self._factorize(hp, jitter=True, want_grad=False, factors=factors)
        self._project(st)
        # every term is linear in the entries of `flat` (trial sums add up over the shards; the replicated log-det term is
        # weighted 1/world), so the raw vector is all-reduced on the device and assembled once
        pieces = [st["sums"][:4], st["sums_b"] if st["sums_b"] is not None else st["sums"][:2] * 0.0, self._info_piece(st)]
        chk = self._theta_check_piece(hp)
        if chk is not None:
            pieces.append(chk)
        flat = self._theta_check(self.shard.allreduce_device(torch.cat(pieces)), chk is not None)
        if np.any(flat[6:] != 0):
            raise np.linalg.LinAlgError("Eigenvalues did not converge")
        f = self.shard.det_fraction()
        return float(-0.5 * self.ntrials_total * f * flat[2] - 0.5 * (flat[0] + flat[4]))

    def loglik_grad(self, hp, factors=None):
        """(loglik, d loglik / d natural parameters) in the order R, ell(s), (ell_t, sigma2_t)..., sig2n[...].
        factors: optional caller-supplied (Qs, ls, Qt, lt) used instead of the eigensolvers (see _given_factors)."""
        if self.use_plan and self.timers is None:
            if factors is not None:
                return self._plan_factors(hp, factors, True)
            ll, g, flag = self.loglik_grad_batch([hp])
            if flag[0] != 0:
                raise np.linalg.LinAlgError("Eigenvalues did not converge")
            return float(ll[0]), g[0]
        nx, nt, ldn, N = self.nx, self.nt, self.ldn, self.ntrials
        st = self._factorize(hp, jitter=True, want_grad=True, factors=factors)
        Bm = self._project(st)
        res = st["sums"]
        stream = self._stream
        f = self.shard.det_fraction()
        ntot = float(self.ntrials_total)
        vec = hp.vector_noise

        # --- segment-weighted SYRKs over the trial batch
        Mt = self._buf("Mt", nt, self.ldt)
        Ms = self._buf("Ms", nx, self.ldx)
        wst = self._buf("ws_syrk_t", max(max(self._q("gpcsd_wsyrk_ws_doubles", o, nx, max(N, 1))
                                             for o in {nt, nt // 2, nt - nt // 2}), 2))
        wss = self._buf("ws_syrk_s", max(max(self._q("gpcsd_wsyrk_ws_doubles", o, nt, max(N, 1)) for o in {nx, max(nx // 2, 1)}), 2))
        Ns = None
        if N > 0:
            if self._t_blocks is not None and self._t_fold():
                # every dKt/dtheta is centrosymmetric too, i.e. block diagonal in the folded basis: only the two diagonal
                # blocks of Mt enter <dL/dKt, dKt/dtheta> (the off-diagonal blocks of the buffer stay zero)
                m = nt // 2
                ms = nt - m
                Mt.zero_()
                self._call("gpcsd_wsyrk", ms, nx, N, self._p(Bm), ldn, nt * ldn, self._p(st["ls"]), self._p(Mt), self.ldt,
                           self._p(wst), stream())
                self._call("gpcsd_wsyrk", m, nx, N, self._p(Bm, ms * ldn), ldn, nt * ldn, self._p(st["ls"]),
                           self._p(Mt, ms * self.ldt + ms), self.ldt, self._p(wst), stream())
            else:
                self._call("gpcsd_wsyrk", nt, nx, N, self._p(Bm), ldn, nt * ldn, self._p(st["ls"]), self._p(Mt), self.ldt,
                           self._p(wst), stream())
            if self._s_blocks is not None:
                # same argument on the channel axis: every dKs/dtheta shares the geometry's reflection symmetry, so only the
                # two diagonal blocks of Ms enter the spatial contractions
                mh = nx // 2
                Ms.zero_()
                self._call("gpcsd_wsyrk", mh, nt, N, self._p(Bm), nt * ldn, ldn, self._p(st["lt"]), self._p(Ms), self.ldx,
                           self._p(wss), stream())
                self._call("gpcsd_wsyrk", mh, nt, N, self._p(Bm, mh * nt * ldn), nt * ldn, ldn, self._p(st["lt"]),
                           self._p(Ms, mh * self.ldx + mh), self.ldx, self._p(wss), stream())
            else:
                self._call("gpcsd_wsyrk", nx, nt, N, self._p(Bm), nt * ldn, ldn, self._p(st["lt"]), self._p(Ms), self.ldx,
                           self._p(wss), stream())
        else:
            Mt.zero_()
            Ms.zero_()
        if vec:
            Ns = self._buf("Ns", nx, self.ldx)
            if N > 0:
                self._call("gpcsd_wsyrk", nx, nt, N, self._p(Bm), nt * ldn, ldn, None, self._p(Ns), self.ldx,
                           self._p(wss), stream())
            else:
                Ns.zero_()

        # --- eigen-basis cores and rotation back:  G = Q X Q^T
        Xs = self._buf("Xs", nx, self.ldx)
        Xt = self._buf("Xt", nt, self.ldt)
        self._call("gpcsd_grad_core", nx, self._p(Ms), self.ldx, self._p(Ns) if vec else None, self.ldx,
                   self._p(st["ls"]), self._p(st["s"]) if vec else None, self._p(st["rowA"]), ntot, 1.0, f,
                   self._p(Xs), self.ldx, stream())
        self._call("gpcsd_grad_core", nt, self._p(Mt), self.ldt, None, 0, self._p(st["lt"]), None, self._p(st["colB"]),
                   ntot, 1.0, f, self._p(Xt), self.ldt, stream())
        Gs = self._rotate(Xs, st["QsT"], nx, self.ldx, "s")
        Gt = self._rotate(Xt, st["QtT"], nt, self.ldt, "t")

        # --- temporal hyperparameters: <Gt, dKt_k/d(ell_k, sigma2_k)>
        ntc, kinds, ells, s2 = self._temporal_spec(hp.temporal)
        wsk = self._buf("ws_ktgrad", self._q("gpcsd_kt_grad_ws_doubles", nt, ntc))
        self._call("gpcsd_kt_grad", nt, self._p(self.t_dev), ntc, kinds, ells, s2, self._p(Gt), self.ldt, self._p(wsk),
                   self._p(res, 8), stream())

        # --- spatial hyperparameters.  With U = A Kg:  dL/dR = 2 <dA, Gs U>,  dL/dell_k = <A, (Gs A) dKg_k>
        G = self.G
        wsd = self._buf("ws_dot", self._q("gpcsd_dot_ws_doubles", nx * G))
        GU = self._buf("GU", nx, G)
        self.gemm(0, nx, G, nx, Gs, self.ldx, 0, st["U"], G, 0, GU, G, 0)
        self._call("gpcsd_dot", nx, G, self._p(st["dA"]), G, self._p(GU), G, self._p(wsd), self._p(res, 4), stream())
        GA = self._buf("GA", nx, G)
        self.gemm(0, nx, G, nx, Gs, self.ldx, 0, st["A"], G, 0, GA, G, 0)
        for k in range(len(hp.ells)):
            dk = self._quad_kernels(hp, deriv_axis=k, tag="_d")
            W = self._apply_quad_kernel(GA, nx, dk, "W")
            self._call("gpcsd_dot", nx, G, self._p(st["A"]), G, self._p(W), G, self._p(wsd), self._p(res, 5 + k), stream())

        # --- one device->host read, then the O(P) host assembly
        pieces = [res[: 8 + 2 * ntc]]
        if vec:
            pieces += [st["rowC"], torch.diagonal(Ns[:, :nx])]
        ninfo = sum(int(i.numel()) for i in st["infos"])
        pieces.append(st["sums_b"] if st["sums_b"] is not None else res[:2] * 0.0)
        pieces.append(self._info_piece(st))
        chk = self._theta_check_piece(hp)
        if chk is not None:
            pieces.append(chk)
        # every term below is linear in the entries of `flat` (trial sums add up over the shards, the replicated
        # trial-independent terms carry the weight f = 1/world): all-reduce the raw vector on the device (NCCL, on this
        # stream), ONE device->host read, then the O(P) host assembly
        flat = self._theta_check(self.shard.allreduce_device(torch.cat([p.reshape(-1) for p in pieces])), chk is not None)
        if np.any(flat[-ninfo:] != 0):
            raise np.linalg.LinAlgError("Eigenvalues did not converge")
        flat = flat[:-ninfo]
        quad, bsq, slogD, srD = flat[0] + flat[-2], flat[1] + flat[-1], flat[2], flat[3]
        ll = -0.5 * ntot * f * slogD - 0.5 * quad
        g = [2.0 * flat[4]] + [flat[5 + k] for k in range(len(hp.ells))] + list(flat[8: 8 + 2 * ntc])
        if vec:
            off = 8 + 2 * ntc
            rowC, dNs = flat[off: off + nx], flat[off + nx: off + 2 * nx]
            g += list(-0.5 * ntot * f * rowC + 0.5 * dNs)
        else:
            g.append(-0.5 * ntot * f * srD + 0.5 * bsq)
        return float(ll), np.array(g, dtype=np.float64)

    def _plan_factors(self, hp, factors, want_grad):
        """Caller-supplied factors through the plan's kernel-level ABI entry (gpcsd_plan_loglik_grad_factors)."""
        Qs, ls, Qt, lt = (np.asarray(a, dtype=np.float64) for a in factors)
        if Qs.shape != (self.nx, self.nx) or Qt.shape != (self.nt, self.nt) or ls.shape != (self.nx,) or lt.shape != (self.nt,):
            raise ValueError("factors must be (Qs (nx,nx), ls (nx,), Qt (nt,nt), lt (nt,))")
        QsT, QtT = self._buf("QT_s_given", self.nx, self.ldx), self._buf("QT_t_given", self.nt, self.ldt)
        QsT[:, :self.nx].copy_(self._dev(Qs.T))
        QtT[:, :self.nt].copy_(self._dev(Qt.T))
        lsd, ltd = self._dev(ls), self._dev(lt)
        pl = self._plan_for(hp, 1)
        out = pl.evaluate_with_factors(self.theta_of(hp), QsT, lsd, QtT, ltd, want_grad)
        if self.shard.enabled and self.shard.world > 1:
            raise RuntimeError("caller-supplied factors are a single-rank (kernel-level) entry")
        P = pl.P
        return float(out[0, 0]), out[0, 1:1 + P].copy()

    def _rotate(self, X, QT, n, ld, tag):
        """G = Q X Q^T from QT = Q^T (row-major):  T1 = X QT ;  G = Q T1."""
        Q = self._buf("Q_" + tag, n, ld)
        self._call("gpcsd_transpose", n, n, self._p(QT), ld, self._p(Q), ld, self._stream())
        T1 = self._buf("rot_tmp_" + tag, n, ld)
        self.gemm(0, n, n, n, X, ld, 0, QT, ld, 0, T1, ld, 0)
        Gm = self._buf("G_" + tag, n, ld)
        self.gemm(0, n, n, n, Q, ld, 0, T1, ld, 0, Gm, ld, 0)
        return Gm

    def shift_objective(self, hp, mu, tau, mutau=0.0, sigtau=10.0, want_grad=True, factors=None):
        """Per-trial evoked-shift objective of auditory_lfp/fit_mean_function.py:311-321 for ALL uploaded trials at once:
            nll_r(tau_r) = 1/2 sum (Qs^T (Y_r - mu(tau_r)) Qt)^2 / D + 1/2 sum_s ((tau_rs - mutau) / sigtau)^2,
            mu(tau) = mu[:, :, 0] + sum_s lerp(mu[:, :, s], t + tau_s)        (scipy interp1d, fill_value="extrapolate")
        with the factors of the fitted model (Ks + JITTER I as at fit_mean_function.py:103).  mu: (nx, nt, nseg+1) host array,
        tau: (ntrials_local, nseg).  The reference runs one scipy L-BFGS-B per trial on joblib workers (:323-328), each
        evaluation a pair of tiny GEMMs; here one evaluation of the whole trial batch is the shifted-residual kernel, the
        Kronecker projection with the fused /D epilogue (the same kernels as loglik), a per-trial column reduction, and --
        for the analytic gradient the reference leaves to finite differences -- the back-projection K^-1 resid = Qs B Qt^T
        and a per-trial contraction with the interpolant's slopes.  Returns (nll (N,), grad (N, nseg) or None)."""
        nx, nt, ldn, N = self.nx, self.nt, self.ldn, self.ntrials
        if self.Y is None:
            raise RuntimeError("no LFP uploaded: call set_lfp first")
        mu = np.ascontiguousarray(np.transpose(np.asarray(mu, dtype=np.float64), (2, 0, 1)))     # [nseg+1][nx][nt]
        nseg = mu.shape[0] - 1
        tau = np.ascontiguousarray(np.asarray(tau, dtype=np.float64).reshape(N, nseg))
        if mu.shape[1:] != (nx, nt):
            raise ValueError("mu must have shape (nx, nt, nseg+1)")
        mu_d, tau_d = self._dev(mu), self._dev(tau)
        Rb = self._buf("shift_resid", nx, nt, ldn)
        if self._y_ready is not None:
            torch.cuda.current_stream(self.device).wait_event(self._y_ready)
        self._call("gpcsd_shift_residual", nx, nt, N, ldn, self._p(self.Y), nseg, self._p(mu_d), self._p(self.t_dev),
                   1 if self.t_uniform else 0, self._p(tau_d), self._p(Rb), self._stream())
        st = self._factorize(hp, jitter=True, want_grad=False, factors=factors, data=Rb)
        Bm = self._project(st)
        nws = self._q("gpcsd_per_trial_ws_doubles", nx, nt, max(N, 1), ldn, max(nseg, 1))
        ws = self._buf("per_trial_ws", nws)
        quad = self._buf("shift_quad", max(N, 1))
        self._call("gpcsd_quad_per_trial", nx, nt, N, ldn, self._p(Bm), self._p(st["rD"]), self.ldt, self._p(ws), self._p(quad),
                   self._stream())
        grad = None
        if want_grad and nseg > 0:
            # V = Qs B Qt^T: rows of Bm are ordered like the rows of QtT (both in the solver's block order), so the plain
            # order-nt products are valid whether or not the projection ran in the folded basis
            Qs, Qt = self._buf("Q_s", nx, self.ldx), self._buf("Q_t", nt, self.ldt)
            self._call("gpcsd_transpose", nx, nx, self._p(st["QsT"]), self.ldx, self._p(Qs), self.ldx, self._stream())
            self._call("gpcsd_transpose", nt, nt, self._p(st["QtT"]), self.ldt, self._p(Qt), self.ldt, self._stream())
            Wb = self._buf("shift_W", nx, nt, ldn)
            self.gemm(0, nt, max(N, 1), nt, Qt, self.ldt, 0, Bm, ldn, nt * ldn, Wb, ldn, nt * ldn, batch=nx)
            V = Rb                                           # the residual is dead after the projection
            self.gemm(0, nx, nt * ldn, nx, Qs, self.ldx, 0, Wb, nt * ldn, 0, V, nt * ldn, 0)
            gd = self._buf("shift_grad", max(N, 1), nseg)
            self._call("gpcsd_shift_grad", nx, nt, N, ldn, self._p(V), nseg, self._p(mu_d), self._p(self.t_dev),
                       1 if self.t_uniform else 0, self._p(tau_d), self._p(ws), self._p(gd), self._stream())
            grad = gd[:N].cpu().numpy() + (tau - mutau) / (sigtau * sigtau)
        self._check_info(st)
        nll = 0.5 * quad[:N].cpu().numpy() + 0.5 * np.sum(np.square((tau - mutau) / sigtau), axis=1)
        return nll, grad

    def predict(self, hp, z, tstar, kind="csd", to_host=True, factors=None):
        """Posterior mean of CSD and/or LFP at (z, t*) per temporal component and summed
        (gpcsd1d.py:248-293 / gpcsd2d.py:289-334), in Kronecker form:
            out_k[:, :, r] = (Kc^T Qs) ((Qs^T Y_r Qt) / D) (Qt^T Kt*_k),   Kt*_k = Kt_k(t*, t)
        No jitter on Ks (gpcsd1d.py:258).  Like the reference (mykron(..).T @ invy), requires len(t*) == len(t).
        Returns {name: array (nz, nt*, ntrials_local)} for name in csd_pred, csd_pred_list[k], ..."""
        nx, nt, ldn, N = self.nx, self.nt, self.ldn, self.ntrials
        z = np.asarray(z, dtype=np.float64)
        nz = z.shape[0]
        ts = np.asarray(tstar, dtype=np.float64).reshape(-1)
        if ts.shape[0] != nt:
            raise ValueError("shapes (%d,%d) and (%d,%d) not aligned: dim 1 != dim 0"
                             % (nz * nt, nx * ts.shape[0], nx * nt, self.ntrials_total))
        st = self._factorize(hp, jitter=False, want_grad=False, factors=factors)
        Bm = self._project(st)
        Qs = self._buf("Q_s", nx, self.ldx)
        Qt = self._buf("Q_t", nt, self.ldt)
        self._call("gpcsd_transpose", nx, nx, self._p(st["QsT"]), self.ldx, self._p(Qs), self.ldx, self._stream())
        self._call("gpcsd_transpose", nt, nt, self._p(st["QtT"]), self.ldt, self._p(Qt), self.ldt, self._stream())
        ts_dev = self._dev(ts)
        G = self.G
        ldz = _even(nx)
        results = {}
        names = [n for n in ("csd", "lfp") if kind in ("both", n)]
        if not names:
            raise ValueError("type must be 'csd', 'lfp' or 'both'")
        for name in names:
            # KcT (nz x nx): transposed cross-covariance
            KcT = self._buf("KcT", nz, ldz)
            if name == "csd":
                # Kphig^T = Kgz^T A^T, Kgz^T[z][g] (covariances.py:58-72 / 188-202)
                KgzT = self._buf("KgzT", nz, G)
                if self.dim == 1:
                    zd = self._dev(z.reshape(-1))
                    self._call("gpcsd_se_matrix", nz, self._p(zd), G, self._p(self.gl_x), float(hp.ells[0]), 1.0, 0,
                               self._p(KgzT), G, self._stream())
                else:
                    zd = self._dev(z.reshape(nz, 2))
                    self._call("gpcsd_se_grid_to_pts", self.G1, self.G2, self._p(self.gl_x1), self._p(self.gl_x2), nz,
                               self._p(zd), float(hp.ells[0]), float(hp.ells[1]), self._p(KgzT), G, self._stream())
                self.gemm(1, nz, nx, G, KgzT, G, 0, st["A"], G, 0, KcT, ldz, 0)
            else:
                # Kphi(xp=z)^T = A_z U^T   (covariances.py:91-95 / 224-231)
                zd = self._dev(z.reshape(-1) if self.dim == 1 else z.reshape(nz, 2))
                Az, _ = self._fwd_weights(zd, nz, hp, False, "z")
                self.gemm(1, nz, nx, G, Az, G, 0, st["U"], G, 0, KcT, ldz, 0)
            Ps = self._buf("Ps", nz, ldz)
            self.gemm(0, nz, nx, nx, KcT, ldz, 0, Qs, self.ldx, 0, Ps, ldz, 0)
            V = self._buf("V", nz, nt, ldn)
            self.gemm(0, nz, nt * ldn, nx, Ps, ldz, 0, Bm, nt * ldn, 0, V, nt * ldn, 0)
            parts = []
            # t* == t on a uniform grid (every caller of the reference): Kt*_k is centrosymmetric like Kt, so in the folded
            # time basis Kt*_k Qt = blockdiag(Cs Us, Ca Ua): two half-order batched GEMMs + one unfold pass per output
            folded = self._t_blocks is not None and self._t_fold() and N > 0 and np.array_equal(ts, self.t_host)
            for k, (knd, ell, s2) in enumerate(hp.temporal):
                ntc1, kinds, ells, s2s = self._temporal_spec([(knd, ell, s2)])
                KtsT = self._buf("KtsT", nt, self.ldt)      # KtsT[j][j'] = k(t_j - t*_j') = Kt*_k[j'][j]
                self._call("gpcsd_kt_build", nt, self._p(self.t_dev), nt, self._p(ts_dev), ntc1, kinds, ells, s2s,
                           self._p(KtsT), self.ldt, self._stream())
                out = torch.empty((nz, nt, ldn), dtype=F64, device=self.device)
                if folded:
                    UsT, lds, UaT, lda = self._t_blocks
                    m = nt // 2
                    ms = nt - m
                    Cs, Ca = self._buf("pred_Cs", ms, lds), self._buf("pred_Ca", max(m, 1), lda)
                    self._call("gpcsd_centro_split", nt, self._p(KtsT), self.ldt, self._p(Cs), lds, self._p(Ca), lda,
                               self._stream())
                    Ts, Ta = self._buf("pred_Ts", ms, lds), self._buf("pred_Ta", max(m, 1), lda)
                    self.gemm(1, ms, ms, ms, Cs, lds, 0, UsT, lds, 0, Ts, lds, 0)          # Cs Us
                    self.gemm(1, m, m, m, Ca, lda, 0, UaT, lda, 0, Ta, lda, 0)             # Ca Ua
                    outf = self._buf("pred_outf", nz, nt, ldn)
                    self._call("gpcsd_dgemm", 0, ms, N, ms, self._p(Ts), lds, 0, self._p(V), ldn, nt * ldn, self._p(outf), ldn,
                               nt * ldn, nz, self._stream(), tag="predict_backproject")
                    self._call("gpcsd_dgemm", 0, m, N, m, self._p(Ta), lda, 0, self._p(V, ms * ldn), ldn, nt * ldn,
                               self._p(outf, ms * ldn), ldn, nt * ldn, nz, self._stream(), tag="predict_backproject")
                    self._call("gpcsd_centro_unfold", nz, nt, ldn, self._p(outf), self._p(out), self._stream())
                else:
                    Tk = self._buf("Tk", nt, self.ldt)           # Tk = Kt*_k^T Qt
                    self.gemm(0, nt, nt, nt, KtsT, self.ldt, 0, Qt, self.ldt, 0, Tk, self.ldt, 0)
                    self.gemm(0, nt, max(N, 1), nt, Tk, self.ldt, 0, V, ldn, nt * ldn, out, ldn, nt * ldn, batch=nz)
                parts.append(self._maybe_download(out, N, to_host))
            tot = torch.empty((nz, nt, ldn), dtype=F64, device=self.device)
            ptrs = (ctypes.c_void_p * len(parts))(*[(p[0] if to_host else p).data_ptr() for p in parts])
            self._call("gpcsd_sum_arrays", nz * nt * ldn, len(parts), ptrs, self._p(tot), self._stream())
            results[name + "_pred"] = self._maybe_download(tot, N, to_host)
            results[name + "_pred_list"] = parts
        self._check_info(st)
        if not to_host:
            return {k: ([p[:, :, :N] for p in v] if isinstance(v, list) else v[:, :, :N]) for k, v in results.items()}
        # every output has been streaming to pinned host memory on the copy stream since it was produced
        self._copy_stream.synchronize()
        return {k: ([h[1].numpy() for h in v] if isinstance(v, list) else v[1].numpy()) for k, v in results.items()}

    def _maybe_download(self, dev, N, to_host):
        """to_host: start the device->host DMA of a finished output on the copy stream (overlaps the GEMMs of the next
        temporal component; pinned buffers come from torch's caching host allocator) and return (device, host)."""
        if not to_host:
            return dev
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        done = torch.cuda.Event()
        done.record(torch.cuda.current_stream(self.device))
        host = torch.empty((dev.shape[0], dev.shape[1], N), dtype=F64, pin_memory=True)
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(done)
            host.copy_(dev[:, :, :N], non_blocking=True)
            dev.record_stream(self._copy_stream)
        return (dev, host)

