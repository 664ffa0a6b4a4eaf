"""TEST INFRASTRUCTURE ONLY -- torch-float64 CPU restatement of the reference's ``obj_fun`` with
reverse-mode autograd through ``torch.linalg.eigh``.

Stand-in for ``autograd.grad(obj_fun)`` (gpcsd1d.py:211 / gpcsd2d.py:250): HIPS autograd is not
installed here, so the reference's own gradient cannot be run (gradient parity UNPINNED, see
gpcsd_oracle.py header).  Reverse-mode AD of the same formula -- including the eigh VJP with
F_ij = 1/(l_j - l_i) that autograd.numpy.linalg.eigh also uses -- is the published algorithm; this
file applies torch's implementation of it to a literal restatement of loglik (gpcsd1d.py:113-128).
Used by tests to cross-check the closed-form gradient in gpcsd_oracle.loglik_and_grad.
"""
import numpy as np
import torch

from . import gpcsd_oracle as O

DT = torch.float64


def _Ks(model, R, ells):
    sp = model.spatial
    if model.dim == 1:
        g = torch.as_tensor(sp.gl_x, dtype=DT)[None, :]
        w = torch.as_tensor(sp.gl_w, dtype=DT)[None, :]
        x = torch.as_tensor(sp.x, dtype=DT)
        q = torch.square((g - x) / R)
        A = w * (torch.sqrt(q + 1.0) - torch.sqrt(q))                       # fwd:16, cov:86-88
        Kg = torch.exp(-0.5 * torch.square((g.T - g) / ells[0]))            # cov:89
    else:
        wq = torch.as_tensor(sp.delta_w(sp.x), dtype=DT)
        wp = torch.as_tensor(sp.w_prod, dtype=DT)[None, :]
        Re = R + model.eps
        A = wp * (torch.log(Re + torch.sqrt(Re ** 2 + wq ** 2))
                  - torch.log(model.eps + torch.sqrt(model.eps ** 2 + wq ** 2)))   # fwd:52, cov:220-221
        g1 = torch.as_tensor(sp.grid1, dtype=DT)
        g2 = torch.as_tensor(sp.grid2, dtype=DT)
        sq1 = torch.square(g1[:, None] - g1[None, :])
        sq2 = torch.square(g2[:, None] - g2[None, :])
        Kg = torch.exp(-0.5 * sq1 / ells[0] ** 2) * torch.exp(-0.5 * sq2 / ells[1] ** 2)  # cov:216
    return (A @ Kg) @ A.T


def loglik_torch(model, lfp, R, ells, temporal, sig2n):
    """loglik as a differentiable function of torch scalars (temporal = [(kind, ell, s2), ...])."""
    Y = torch.as_tensor(np.atleast_3d(lfp), dtype=DT)
    nx, nt, N = Y.shape
    Ks = _Ks(model, R, ells) + model.jitter * torch.eye(nx, dtype=DT)
    t = torch.as_tensor(np.asarray(model.t).reshape(-1, 1), dtype=DT)
    dist = t - t.T
    Kt = torch.zeros((nt, nt), dtype=DT)
    for kind, ell, s2 in temporal:
        if kind == O.KIND_SE:
            Kt = Kt + s2 * torch.exp(-0.5 * torch.square(dist) / torch.square(ell))
        else:
            Kt = Kt + s2 * torch.exp(-torch.sqrt(torch.square(dist)) / ell)
    lt, Qt = torch.linalg.eigh(Kt)
    ls, Qs = torch.linalg.eigh(Ks)
    if sig2n.ndim == 0:
        nvec = sig2n * torch.ones(nx * nt, dtype=DT)
    else:
        nvec = torch.repeat_interleave(sig2n, nt)                            # util:57
    D = torch.repeat_interleave(ls, nt) * ls.new_tensor(1.0) * lt.repeat(nx) + nvec
    A = torch.einsum("ia,ijr,jb->abr", Qs, Y, Qt).reshape(nx * nt, N)
    return -0.5 * N * torch.sum(torch.log(D)) - 0.5 * torch.sum(torch.square(A) / D[:, None])


def loglik_and_grad_torch(model, lfp):
    """(loglik, gradient w.r.t. natural parameters) in the same order as gpcsd_oracle.loglik_and_grad."""
    leaves = []

    def leaf(v):
        x = torch.tensor(np.asarray(v, dtype=np.float64), dtype=DT, requires_grad=True)
        leaves.append(x)
        return x

    R = leaf(model.R)
    ells = [leaf(e) for e in model.ells]
    temporal = [(k, leaf(e), leaf(s)) for k, e, s in model.temporal]
    s2n = leaf(model.sig2n)
    ll = loglik_torch(model, lfp, R, ells, temporal, s2n)
    grads = torch.autograd.grad(ll, leaves)
    return float(ll.detach()), np.concatenate([np.atleast_1d(g.numpy()) for g in grads])
