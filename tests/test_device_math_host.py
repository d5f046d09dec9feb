"""fp32 device math (csrc/aep_math.cuh) compiled for the HOST through tests/cpu_math_harness.cpp and checked against
the fp64 oracle.  CPU only; pins B-spline, SVD, stress and return-mapping numerics before any GPU time is spent.
The harness is test infrastructure: the product never loads it."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle_py as op

HERE = os.path.dirname(os.path.abspath(__file__))
fp = C.POINTER(C.c_float); dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def H(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("harness") / "libmath_host.so")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", so, os.path.join(HERE, "cpu_math_harness.cpp")])
    h = C.CDLL(so)
    h.h_bspline4.argtypes = [C.c_float, fp, fp]; h.h_bspline_lane.argtypes = [C.c_float, C.c_int, fp, fp]
    h.h_stress.argtypes = [C.c_int, C.c_double, C.c_double, fp, fp, C.c_float, C.c_float, fp]
    h.h_return_map.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, fp, fp, fp, fp]
    return h


def P(a): return a.ctypes.data_as(fp)
def D(a): return a.ctypes.data_as(dp)
def f32(a): return np.ascontiguousarray(a, np.float32)


def test_bspline_forms_match_reference(H):
    """thread form (4 nodes) and lane form (one node) both equal interpolation.cpp:9-33 at u = f + 1 - o."""
    rng = np.random.default_rng(0)
    for f in list(rng.random(300)) + [0.0, 0.99999994, 0.5, 1e-7]:
        N = np.zeros(4, np.float32); Dv = np.zeros(4, np.float32); H.h_bspline4(C.c_float(f), P(N), P(Dv))
        for o in range(4):
            u = float(np.float32(f)) + 1 - o
            n1 = C.c_float(); d1 = C.c_float(); H.h_bspline_lane(C.c_float(f), o, C.byref(n1), C.byref(d1))
            assert abs(N[o] - op.cubic_bspline(u)) < 2e-7 and abs(Dv[o] - op.dcubic_bspline(u)) < 3e-7
            assert abs(n1.value - N[o]) < 1e-7 and abs(d1.value - Dv[o]) < 1e-7


def test_svd3_fp32(H):
    rng = np.random.default_rng(1)
    for k in range(4000):
        kind = k % 4
        if kind == 0: F = np.eye(3) + 1e-3 * rng.standard_normal((3, 3))
        elif kind == 1: F = np.eye(3) + 0.3 * rng.standard_normal((3, 3))
        elif kind == 2: F = rng.standard_normal((3, 3))
        else:
            Q, _ = np.linalg.qr(rng.standard_normal((3, 3))); F = Q @ np.diag([1.0, 1.0 + 1e-6, 1.0 - 1e-6]) @ Q.T
        F32 = f32(F); U = np.zeros(9, np.float32); S = np.zeros(3, np.float32); V = np.zeros(9, np.float32)
        H.h_svd3(P(F32.ravel()), P(U), P(S), P(V))
        U = U.reshape(3, 3).astype(np.float64); V = V.reshape(3, 3).astype(np.float64); S = S.astype(np.float64)
        scale = max(1.0, np.abs(F32).max())
        assert (S >= 0).all()                                             # Eigen contract: non-negative singular values
        assert np.abs(U @ np.diag(S) @ V.T - F32).max() < 2e-6 * scale
        assert np.abs(U.T @ U - np.eye(3)).max() < 2e-6 and np.abs(V.T @ V - np.eye(3)).max() < 2e-6
        assert np.abs(np.sort(S)[::-1] - np.linalg.svd(F32.astype(np.float64), compute_uv=False)).max() < 1e-6 * scale


@pytest.mark.parametrize("mat,E,nu", [(1, 3.537e5, 0.3), (0, 1.4e5, 0.2)])
def test_stress_and_return_map_vs_oracle(H, mat, E, nu):
    """fp32 stress error is bounded by eps_fp32 / strain (F is stored in fp32); return mapping to ~1e-6 absolute."""
    L = op.lib()
    L.orc_particle_stress.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, dp, dp, dp, C.c_double, dp]
    L.orc_particle_return_map.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, dp, dp, dp, dp]
    rng = np.random.default_rng(2)
    cm = lambda M: np.ascontiguousarray(M.T).ravel()
    for strain in (1e-3, 1e-2, 5e-2, 0.2):
        worst = 0.0; amax = 0.0
        for _ in range(300):
            FE = f32(np.eye(3) + strain * rng.standard_normal((3, 3))); Fh = f32(FE + 0.3 * strain * rng.standard_normal((3, 3)))
            FPm = f32(np.eye(3) + 0.02 * rng.standard_normal((3, 3))); vol = 1e-6; q0 = float(np.float32(abs(rng.standard_normal()) * 0.3))
            Jp = np.float32(np.linalg.det(FPm.astype(np.float64)))
            A = np.zeros(9, np.float32); H.h_stress(mat, E, nu, P(Fh.ravel()), P(FE.ravel()), C.c_float(vol), C.c_float(Jp), P(A))
            A64 = np.zeros(9)
            L.orc_particle_stress(mat, E, nu, 10.0, D(cm(Fh.astype(np.float64))), D(cm(FE.astype(np.float64))), D(cm(FPm.astype(np.float64))), vol, D(A64))
            A64 = A64.reshape(3, 3).T
            worst = max(worst, np.abs(A.reshape(3, 3) - A64).max()); amax = max(amax, np.abs(A64).max())
            FEo = np.zeros(9, np.float32); FPo = FPm.copy().ravel(); q = C.c_float(q0)
            H.h_return_map(mat, E, nu, 2.5e-2, 7.5e-3, P(Fh.ravel()), P(FEo), P(FPo), C.byref(q))
            FE9 = np.zeros(9); FP9 = cm(FPm.astype(np.float64)).copy(); q64 = C.c_double(q0)
            L.orc_particle_return_map(mat, E, nu, 2.5e-2, 7.5e-3, D(cm(Fh.astype(np.float64))), D(FE9), D(FP9), C.byref(q64))
            assert np.abs(FEo.reshape(3, 3) - FE9.reshape(3, 3).T).max() < 3e-6
            assert np.abs(FPo.reshape(3, 3) - FP9.reshape(3, 3).T).max() < 3e-6
            assert abs(q.value - q64.value) < 3e-6
        assert worst / amax < 3e-7 / strain + 1e-6, (strain, worst / amax)


def _weights64(f):
    """cubic B-spline values of the 4 stencil nodes of a particle at cell fraction f, fp64, straight from interpolation.cpp:9-16
    via the oracle: node o sits at signed distance u = f + 1 - o (in cells)."""
    return np.array([op.cubic_bspline(f + 1 - o) for o in range(4)])


def _dweights64(f, h):
    """d w / d x_p per axis = N'(u) / h (HybridSolver.cpp:48-57, interpolation.cpp:18-33)"""
    return np.array([op.dcubic_bspline(f + 1 - o) for o in range(4)]) / h


def test_p2g_scatter_rows_match_reference_formula(H):
    """The packed phase-B code of k_p2g (aep_scatter.cuh: record + 16 row accumulations, run here on the host) against
    HybridSolver.cpp:113-231 written out directly: m_i += w m, p_i += w m (v + (3/h_min^2) B (x_i - x_p))."""
    H.h_p2g_scatter.argtypes = [fp, fp, fp, C.c_float, fp, fp]
    rng = np.random.default_rng(5)
    for trial in range(200):
        f = f32(rng.random(3) * 0.999); v = f32(rng.standard_normal(3)); B = f32(0.1 * rng.standard_normal((3, 3)))
        h = f32(np.array([1 / 64, 1 / 48, 1 / 80]) if trial % 2 else np.full(3, 1 / 128)); m = np.float32(abs(rng.standard_normal()) + 0.1)
        out = np.zeros(64 * 4, np.float32)
        H.h_p2g_scatter(P(f), P(v), P(B.ravel()), C.c_float(m), P(h), P(out))
        out = out.reshape(4, 4, 4, 4)                                       # [k][j][i][c]
        f64 = f.astype(np.float64); h64 = h.astype(np.float64); B64 = B.astype(np.float64); v64 = v.astype(np.float64); m64 = float(m)
        W = [_weights64(f64[a]) for a in range(3)]
        apic = 3.0 / h64.min() ** 2
        ref = np.zeros((4, 4, 4, 4))
        for k in range(4):
            for j in range(4):
                for i in range(4):
                    w = W[0][i] * W[1][j] * W[2][k]
                    r = h64 * (np.array([i, j, k]) - 1.0 - f64)             # x_i - x_p
                    ref[k, j, i, 0] = w * m64
                    ref[k, j, i, 1:] = w * m64 * (v64 + apic * (B64 @ r))
        assert np.abs(out[..., 0] - ref[..., 0]).max() < 2e-6 * np.abs(ref[..., 0]).max()
        assert np.abs(out[..., 1:] - ref[..., 1:]).max() < 3e-6 * np.abs(ref[..., 1:]).max()
        assert abs(out[..., 0].sum() - m64) < 2e-6 * m64                    # partition of unity through the packed path


def test_force_scatter_rows_match_reference_formula(H):
    """The packed phase-B code of k_forces against HybridSolver.cpp:356-366: f_i += A grad w_i."""
    H.h_frc_scatter.argtypes = [fp, fp, fp, fp]
    rng = np.random.default_rng(6)
    for trial in range(200):
        f = f32(rng.random(3) * 0.999); A = f32(rng.standard_normal((3, 3)))
        h = f32(np.array([1 / 64, 1 / 48, 1 / 80]) if trial % 2 else np.full(3, 1 / 128))
        out = np.zeros(64 * 4, np.float32)
        H.h_frc_scatter(P(f), P(A.ravel()), P(h), P(out))
        out = out.reshape(4, 4, 4, 4)
        f64 = f.astype(np.float64); h64 = h.astype(np.float64); A64 = A.astype(np.float64)
        N = [_weights64(f64[a]) for a in range(3)]; Dn = [_dweights64(f64[a], h64[a]) for a in range(3)]
        ref = np.zeros((4, 4, 4, 3))
        for k in range(4):
            for j in range(4):
                for i in range(4):
                    g = np.array([Dn[0][i] * N[1][j] * N[2][k], N[0][i] * Dn[1][j] * N[2][k], N[0][i] * N[1][j] * Dn[2][k]])
                    ref[k, j, i] = A64 @ g
        assert np.abs(out[..., :3] - ref).max() < 3e-6 * np.abs(ref).max()
        assert np.abs(out[..., 3]).max() == 0.0                             # the fourth lane of the force record stays zero
        assert np.abs(out[..., :3].sum(axis=(0, 1, 2))).max() < 2e-5 * np.abs(ref).max()   # sum_i grad w_i = 0: no net force from one particle
