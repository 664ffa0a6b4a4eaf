"""Numerics of the in-house divide-and-conquer tridiagonal eigensolver, checked WITHOUT a GPU: the numerical core
(gpcsd_b200/csrc/dc_core.h: deflation, secular roots, Gu-Eisenstat vectors) is the same source the CUDA kernel compiles;
tests/dc_host_harness.cpp drives it sequentially and this test compares with LAPACK on the matrix families that matter
(GP kernels with a numerically degenerate tail) and on classic hard cases (Wilkinson, glued, graded, repeated)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import scipy.linalg

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def dc_host(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("dc") / "libdc_host.so")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-o", out, os.path.join(HERE, "dc_host_harness.cpp")], check=True)
    lib = ctypes.CDLL(out)
    lib.dc_host_eig.restype = ctypes.c_long

    def run(d, e):
        n = len(d)
        W, QT, ee = np.zeros(n), np.zeros((n, n)), np.zeros(n)
        ee[1:] = e
        P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        lib.dc_host_eig(n, P(np.ascontiguousarray(d, dtype=np.float64)), P(ee), P(W), P(QT))
        return W, QT
    return run


def _tridiag_of(M):
    H = scipy.linalg.hessenberg(M)
    return np.diag(H).copy(), np.diag(H, -1).copy()


def _families(n, rng):
    t = np.arange(n) * 0.7
    dd = t[:, None] - t[None, :]
    yield "random", rng.standard_normal(n), rng.standard_normal(n - 1)
    yield "se+matern", *_tridiag_of(0.5 * np.exp(-0.5 * dd ** 2 / 30.0) + 0.2 * np.exp(-np.abs(dd) / 4.0))
    yield "se (degenerate tail)", *_tridiag_of(3.0 * np.exp(-0.5 * dd ** 2 / 30.0))
    yield "identity", np.ones(n), np.zeros(n - 1)
    yield "1-2-1", 2 * np.ones(n), -np.ones(n - 1)
    yield "wilkinson", np.abs(np.arange(n) - (n - 1) / 2), np.ones(n - 1)
    yield "glued", np.tile(np.arange(1, 6.0), (n + 4) // 5)[:n], np.where(np.arange(n - 1) % 5 == 4, 1e-9, 1.0)
    yield "tiny off-diagonal", rng.standard_normal(n), 1e-18 * rng.standard_normal(n - 1)
    yield "scaled 1e-200", 1e-200 * rng.standard_normal(n), 1e-200 * rng.standard_normal(n - 1)
    yield "graded", 10.0 ** (-np.arange(n) * 16.0 / n), 10.0 ** (-np.arange(1, n) * 16.0 / n)
    yield "negative off-diagonals", rng.standard_normal(n), -np.abs(rng.standard_normal(n - 1))


@pytest.mark.parametrize("n", [2, 3, 5, 8, 17, 24, 64, 125, 192, 250, 256])
def test_dc_core_matches_lapack(dc_host, n):
    rng = np.random.default_rng(n)
    for name, d, e in _families(n, rng):
        T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
        W, QT = dc_host(d, e)
        lam = scipy.linalg.eigvalsh_tridiagonal(d, e)
        sc = max(np.max(np.abs(T)), 1e-300)
        Q = QT.T
        assert np.all(np.diff(W) >= 0), name
        assert np.max(np.abs(W - lam)) <= 5e-14 * sc, name
        assert np.max(np.abs(Q.T @ Q - np.eye(n))) <= 2e-14, name
        assert np.max(np.abs(T @ Q - Q * W)) <= 2e-14 * sc, name
