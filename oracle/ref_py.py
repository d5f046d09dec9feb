"""ctypes binding of oracle/_ref/libaep_ref.so -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

libaep_ref.so is the reference's own, unmodified C++ (HybridSolver.cpp, ParticleSystem.cpp, RegularGrid.cpp,
LagrangianMesh.cpp, geometry.cpp, interpolation.cpp, LevelSet.cpp compiled where they lie under /root/reference) linked
against the MiniEigen / viewer stand-ins of oracle/ref_shim and the C entry points of oracle/ref_driver.cpp.  It exists to
pin the oracle: `Reference` has the interface of `oracle_py.Oracle`, so the same scene runs through both.

The library can only be BUILT where /root/reference exists (the dev container); the built file travels to the GPU box.
`available()` says whether it is there; the committed fixtures tests/golden/ref_*.npz (oracle/make_ref_golden.py) carry its
outputs to places where it is not.  Importers allowed: tests/, bench.py (cpu_baseline leg and --impl reference).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from anisotropicelastoplasticity_b200.scenes import LS_GROUND, LS_NONE, LS_WALL2GROUND, Scene
from oracle import oracle_py
from oracle.oracle_py import Oracle, _dp, _p

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libaep_ref.so")
REF_SRC = "/root/reference/AnisotropicElastoplasticity"
_LIB = None


def available() -> bool:
    """True when the library is built, or can be built because the reference sources are here."""
    return os.path.exists(_SO) or os.path.exists(os.path.join(REF_SRC, "HybridSolver.cpp"))


def build() -> str:
    if os.path.exists(os.path.join(REF_SRC, "HybridSolver.cpp")):
        subprocess.check_call(["make", "-C", _HERE, "-s", "_ref/libaep_ref.so"])
    if not os.path.exists(_SO):
        raise RuntimeError("oracle/_ref/libaep_ref.so is not built and /root/reference is not present to build it from")
    return _SO


class _Renamed:
    """Lets Oracle's method bodies (written against orc_*) drive the ref_* entry points."""

    def __init__(self, cdll):
        self._c = cdll

    def __getattr__(self, name):
        return getattr(self._c, "ref_" + name[4:] if name.startswith("orc_") else name)


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.ref_create.restype = C.c_void_p
        L.ref_create.argtypes = [_dp, _dp, C.POINTER(C.c_int)]
        for name in ("ref_substep", "ref_get_dt", "ref_get_time", "ref_cfl_condition", "ref_cubic_bspline", "ref_dcubic_bspline",
                     "ref_clamp", "ref_ls_phi"):
            getattr(L, name).restype = C.c_double
        L.ref_cubic_bspline.argtypes = [C.c_double]; L.ref_dcubic_bspline.argtypes = [C.c_double]
        L.ref_clamp.argtypes = [C.c_double] * 3
        L.ref_get_frame.restype = C.c_int; L.ref_get_threads.restype = C.c_int
        L.ref_last_error.restype = C.c_char_p
        L.ref_factory.restype = C.c_long
        _LIB = _Renamed(L)
    return _LIB


class Reference(Oracle):
    """The reference's own classes behind the Oracle interface (stage calls, substep(), particles(), grid(), mesh())."""

    _load = staticmethod(lambda: lib())

    def __init__(self, scene: Scene, rate_floor: float = 3e2):
        super().__init__(scene, threads=1, rate_floor=rate_floor)
        err = self.L.ref_last_error(self.h)
        if err:                                                  # e.g. a parameter the reference hard-codes was asked to differ
            raise RuntimeError(err.decode())

    def _set_levelset(self, scene):
        ls = scene.levelset
        if ls.kind == LS_NONE:
            return
        par = np.ascontiguousarray(ls.params, np.float64)
        if ls.kind in (LS_GROUND, LS_WALL2GROUND):              # the two colliders the reference ships (LevelSet.h:8-12)
            self.L.ref_set_levelset(self.h, C.c_int(ls.kind), _p(par))
            return
        # colliders the reference does not have (sphere, box): hand its collision code phi<=0 flags and normals sampled at the
        # nodes -- the only places HybridSolver.cpp:473-482 evaluates them
        inside, nrm = sample_levelset(scene)
        self.L.ref_set_levelset_samples(self.h, inside.ctypes.data_as(C.POINTER(C.c_uint8)), _p(nrm))

    def solve(self, maxt: float, workdir: str, cfl: float = 0.3):
        """HybridSolver::solve(cfl, maxt, 0.95) itself (sand, rate floor 300), OBJ frames written under workdir."""
        if self.L.ref_solve(self.h, C.c_double(cfl), C.c_double(maxt), os.fsencode(workdir)) != 0:
            raise RuntimeError(self.L.ref_last_error(self.h).decode())


def sample_levelset(scene: Scene):
    """(inside uint8[Ng], normals float64[3*Ng] plane-major) of the scene's analytic collider at the grid nodes."""
    g = scene.grid; par = np.ascontiguousarray(scene.levelset.params, np.float64)
    res = np.asarray(g.res); h = g.h; mn = np.asarray(g.mn, np.float64)
    O = oracle_py.lib()
    O.orc_ls_phi.restype = C.c_double
    ng = g.n_nodes
    inside = np.zeros(ng, np.uint8); nrm = np.zeros(3 * ng)
    x = np.empty(3); n = np.empty(3)
    idx = 0
    for k in range(res[2]):
        for j in range(res[1]):
            for i in range(res[0]):
                x[:] = (mn[0] + i * h[0], mn[1] + j * h[1], mn[2] + k * h[2])      # HybridSolver.cpp:473-476
                if O.orc_ls_phi(C.c_int(scene.levelset.kind), _p(par), _p(x)) <= 0.0:
                    inside[idx] = 1
                    n[:] = (0.0, 0.0, 1.0)
                    O.orc_ls_normal(C.c_int(scene.levelset.kind), _p(par), _p(x), _p(n))
                    nrm[idx], nrm[ng + idx], nrm[2 * ng + idx] = n
                idx += 1
    return inside, nrm


# ---- the reference's scalar kernels
def cubic_bspline(x): return lib().ref_cubic_bspline(C.c_double(x))
def dcubic_bspline(x): return lib().ref_dcubic_bspline(C.c_double(x))
def clamp(x, lo, hi): return lib().ref_clamp(C.c_double(x), C.c_double(lo), C.c_double(hi))


def gram_schmidt(A):
    Ac = np.ascontiguousarray(np.asarray(A, np.float64).T).ravel(); Q = np.empty(9); R = np.empty(9)
    lib().ref_gram_schmidt(_p(Ac), _p(Q), _p(R))
    return Q.reshape(3, 3).T.copy(), R.reshape(3, 3).T.copy()


def svd3(F):
    Fc = np.ascontiguousarray(np.asarray(F, np.float64).T).ravel(); U = np.empty(9); s = np.empty(3); V = np.empty(9)
    lib().ref_svd3(_p(Fc), _p(U), _p(s), _p(V))
    return U.reshape(3, 3).T.copy(), s, V.reshape(3, 3).T.copy()


def svd2(A):
    Ac = np.ascontiguousarray(np.asarray(A, np.float64).T).ravel(); U = np.empty(4); s = np.empty(2); V = np.empty(4)
    lib().ref_svd2(_p(Ac), _p(U), _p(s), _p(V))
    return U.reshape(2, 2).T.copy(), s, V.reshape(2, 2).T.copy()


def ls_phi(kind, params, x):
    return lib().ref_ls_phi(C.c_int(kind), _p(np.ascontiguousarray(params, np.float64)), _p(np.ascontiguousarray(x, np.float64)))


def ls_normal(kind, params, x):
    n = np.empty(3)
    lib().ref_ls_normal(C.c_int(kind), _p(np.ascontiguousarray(params, np.float64)), _p(np.ascontiguousarray(x, np.float64)), _p(n))
    return n


def factory(which, a, b, r, height, n):
    """ParticleSystem::{SnowBall,SandBall,SandBlock,SandCylinder} -> (positions (n,3), masses, [E, nu, thetaC, thetaS, friction])."""
    a = np.ascontiguousarray(a, np.float64); b = np.ascontiguousarray(b, np.float64)
    x = np.empty(3 * n); m = np.empty(n); c = np.empty(5)
    got = lib().ref_factory(C.c_int(which), _p(a), _p(b), C.c_double(r), C.c_double(height), C.c_int(n), _p(x), _p(m), _p(c))
    assert got == n
    return x.reshape(3, n).T.copy(), m, c


def obj_mesh(path, density, thickness, E, nu, shear, stiff, angle_deg):
    """LagrangianMesh::ObjMesh -> dict of what the loader derives (LagrangianMesh.cpp:197-352)."""
    L = lib(); nn = (C.c_long * 2)()
    args = [os.fsencode(path)] + [C.c_double(v) for v in (density, thickness, E, nu, shear, stiff, angle_deg)]
    L.ref_obj_mesh(*args, nn, None, None, None, None, None, None, None, None)
    nv, nf = nn[0], nn[1]
    vx = np.empty(3 * nv); faces = np.empty(3 * nf, np.int32); vm = np.empty(nv); vvol = np.empty(nv); em = np.empty(nf); evol = np.empty(nf)
    eD = np.empty(9 * nf); c = np.empty(3)
    L.ref_obj_mesh(*args, nn, _p(vx), faces.ctypes.data_as(C.POINTER(C.c_int)), _p(vm), _p(vvol), _p(em), _p(evol), _p(eD), _p(c))
    return dict(vx=vx.reshape(3, nv).T.copy(), faces=faces.reshape(3, nf).T.copy(), vm=vm, vvol=vvol, em=em, evol=evol,
                eD=np.stack([eD[3 * nf * a:3 * nf * (a + 1)].reshape(3, nf).T for a in range(3)]), mu=c[0], lam=c[1], fric=c[2])
