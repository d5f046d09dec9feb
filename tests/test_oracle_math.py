"""Known-answer tests of the oracle's math kernels (derivable from the maths).
CPU only.  Reference: interpolation.cpp:9-49, geometry.cpp:31-73, Eigen::JacobiSVD contract, LevelSet.cpp:8-42."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle_py as op


def test_bspline_known_values():
    # N(0)=2/3, N(1)=1/6, N(2)=0, N(0.5)=23/48 ; N'(0)=0, N'(1)=-1/2, N'(-1)=+1/2, N'(2)=0
    assert op.cubic_bspline(0.0) == pytest.approx(2.0 / 3.0, abs=1e-15)
    assert op.cubic_bspline(1.0) == pytest.approx(1.0 / 6.0, abs=1e-15)
    assert op.cubic_bspline(-1.0) == pytest.approx(1.0 / 6.0, abs=1e-15)
    assert op.cubic_bspline(2.0) == 0.0 and op.cubic_bspline(-2.5) == 0.0
    assert op.cubic_bspline(0.5) == pytest.approx(23.0 / 48.0, abs=1e-15)
    assert op.dcubic_bspline(0.0) == 0.0
    assert op.dcubic_bspline(1.0) == pytest.approx(-0.5, abs=1e-15)
    assert op.dcubic_bspline(-1.0) == pytest.approx(0.5, abs=1e-15)
    assert op.dcubic_bspline(2.0) == 0.0 and op.dcubic_bspline(-2.0) == pytest.approx(0.0, abs=1e-15)


def test_bspline_partition_of_unity_and_derivative():
    rng = np.random.default_rng(0)
    for f in rng.random(200):
        u = np.array([f + 1 - o for o in range(4)])
        n = np.array([op.cubic_bspline(x) for x in u]); d = np.array([op.dcubic_bspline(x) for x in u])
        assert n.sum() == pytest.approx(1.0, abs=1e-14)            # sum w = 1
        assert d.sum() == pytest.approx(0.0, abs=1e-14)            # sum grad w = 0
        assert (n * (-u)).sum() == pytest.approx(0.0, abs=1e-14)   # linear reproduction: sum w (x_i - x_p) = 0
        eps = 1e-6                                                 # derivative consistent with the value
        for x in u:
            fd = (op.cubic_bspline(x + eps) - op.cubic_bspline(x - eps)) / (2 * eps)
            assert op.dcubic_bspline(x) == pytest.approx(fd, abs=1e-8)


def test_svd3_contract():
    rng = np.random.default_rng(1)
    for k in range(300):
        F = np.eye(3) + [1e-6, 1e-2, 1.0][k % 3] * rng.standard_normal((3, 3))
        U, s, V = op.svd3(F)
        assert np.allclose(U @ np.diag(s) @ V.T, F, atol=1e-13)
        assert np.allclose(U.T @ U, np.eye(3), atol=1e-13) and np.allclose(V.T @ V, np.eye(3), atol=1e-13)
        assert s[0] >= s[1] >= s[2] >= 0.0                          # Eigen::JacobiSVD: non-negative, sorted descending
        assert np.allclose(s, np.linalg.svd(F, compute_uv=False), atol=1e-13)
    # inverted element: sigma stays >= 0 (log(sigma) at HybridSolver.cpp:330 needs that), reflection goes into U or V
    F = np.diag([1.0, 1.0, -0.5]); U, s, V = op.svd3(F)
    assert (s >= 0).all() and np.allclose(U @ np.diag(s) @ V.T, F, atol=1e-14)


def test_svd2_polar():
    rng = np.random.default_rng(2)
    for _ in range(200):
        A = np.triu(rng.standard_normal((2, 2))); A[0, 0] = abs(A[0, 0]) + 0.1; A[1, 1] = abs(A[1, 1]) + 0.1
        U, s, V = op.svd2(A)
        assert np.allclose(U @ np.diag(s) @ V.T, A, atol=1e-13) and s[0] >= s[1] >= 0
        R = U @ V.T
        assert np.allclose(R.T @ R, np.eye(2), atol=1e-13) and np.linalg.det(R) == pytest.approx(1.0, abs=1e-12)


def test_gram_schmidt_qr():
    rng = np.random.default_rng(3)
    for _ in range(200):
        A = np.eye(3) + 0.4 * rng.standard_normal((3, 3))
        Q, R = op.gram_schmidt(A)
        assert np.allclose(Q @ R, A, atol=1e-13) and np.allclose(Q.T @ Q, np.eye(3), atol=1e-12)
        assert np.allclose(np.tril(R, -1), 0) and (np.diag(R) >= 0).all()        # geometry.cpp:31-62: norms on the diagonal


def test_levelsets():
    L = op.lib(); dp = C.POINTER(C.c_double)
    L.orc_ls_phi.argtypes = [C.c_int, dp, dp]; L.orc_ls_normal.argtypes = [C.c_int, dp, dp, dp]
    def phi(kind, P, x):
        P = np.array(P + [0.0] * (8 - len(P))); x = np.array(x, float)
        return L.orc_ls_phi(kind, P.ctypes.data_as(dp), x.ctypes.data_as(dp))
    def nrm(kind, P, x):
        P = np.array(P + [0.0] * (8 - len(P))); x = np.array(x, float); n = np.zeros(3)
        L.orc_ls_normal(kind, P.ctypes.data_as(dp), x.ctypes.data_as(dp), n.ctypes.data_as(dp)); return n
    assert phi(1, [0.25], [0.3, 0.4, 0.75]) == pytest.approx(0.5)                 # LevelSet.cpp:8-11
    assert (nrm(1, [0.25], [0, 0, 0]) == [0, 0, 1]).all()
    assert phi(2, [1.0, 1.0, 0.0], [0.9, 0.2, 0.5]) == pytest.approx(0.1)         # LevelSet.cpp:18-21 min(z-g, wx-x, wy-y)
    assert (nrm(2, [1.0, 1.0, 0.0], [0.9, 0.2, 0.5]) == [-1, 0, 0]).all()         # closest: x wall
    assert (nrm(2, [1.0, 1.0, 0.0], [0.2, 0.95, 0.5]) == [0, -1, 0]).all()
    assert (nrm(2, [1.0, 1.0, 0.0], [0.2, 0.2, 0.01]) == [0, 0, 1]).all()
    assert phi(3, [0.5, 0.5, 0.2, 0.1, 0.0], [0.5, 0.5, 0.25]) == pytest.approx(-0.05)
    assert np.allclose(nrm(3, [0.5, 0.5, 0.2, 0.1, 0.0], [0.5, 0.5, 0.25]), [0, 0, 1])
    assert phi(4, [0, 0, 0, 1, 1, 1], [0.1, 0.5, 0.5]) == pytest.approx(0.1)
    assert (nrm(4, [0, 0, 0, 1, 1, 1], [0.95, 0.5, 0.5]) == [-1, 0, 0]).all()
