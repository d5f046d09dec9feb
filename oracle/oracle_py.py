"""ctypes binding of oracle/libmpm_oracle.so -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Importers allowed: tests/, __graft_entry__.smoke(), bench.py (cpu_baseline leg and --impl reference).
The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from anisotropicelastoplasticity_b200.scenes import (Scene, colmajor, from_colmajor, mats_colmajor, mats_from_colmajor)

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_dp = C.POINTER(C.c_double)


def build(force=False):
    so = os.path.join(_HERE, "libmpm_oracle.so")
    src = os.path.join(_HERE, "mpm_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [_dp, _dp, C.POINTER(C.c_int)]
        for name in ("orc_substep", "orc_get_dt", "orc_get_time", "orc_cfl_condition", "orc_cubic_bspline",
                     "orc_dcubic_bspline", "orc_ls_phi"):
            getattr(L, name).restype = C.c_double
        L.orc_cubic_bspline.argtypes = [C.c_double]; L.orc_dcubic_bspline.argtypes = [C.c_double]
        L.orc_get_frame.restype = C.c_int; L.orc_get_threads.restype = C.c_int
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


class Oracle:
    """One simulation on the CPU oracle; method names mirror the C ABI of the engine (include/aep_b200.h)."""

    _load = staticmethod(lambda: lib())        # oracle/ref_py.py swaps in the reference's own code behind the same entry points

    def __init__(self, scene: Scene, threads: int = 1, rate_floor: float = 3e2):
        L = self._load(); self.L = L
        g = scene.grid
        mn = np.asarray(g.mn, np.float64); mx = np.asarray(g.mx, np.float64); res = np.asarray(g.res, np.int32)
        self.h = C.c_void_p(L.orc_create(_p(mn), _p(mx), res.ctypes.data_as(C.POINTER(C.c_int))))
        self.ng = g.n_nodes; self.np = 0; self.nv = 0; self.nf = 0
        sand_h = np.array([35.0, 9.0, 0.2, 10.0])
        L.orc_set_params(self.h, C.c_int(scene.material), C.c_double(scene.cfl), C.c_double(9.8), C.c_double(0.2),
                         C.c_double(10.0), _p(sand_h), C.c_double(rate_floor), C.c_double(1.0 / 60.0))
        L.orc_set_threads(self.h, C.c_int(threads))
        if scene.particles is not None:
            p = scene.particles; self.np = p.n
            arrs = [colmajor(p.x), colmajor(p.v), colmajor(p.B[:, 0, :]), colmajor(p.B[:, 1, :]), colmajor(p.B[:, 2, :]),
                    mats_colmajor(p.FE), mats_colmajor(p.FP), np.ascontiguousarray(p.m, np.float64),
                    np.ascontiguousarray(p.vol, np.float64), np.ascontiguousarray(p.q, np.float64)]
            L.orc_set_particles(self.h, C.c_long(p.n), *[_p(a) for a in arrs], C.c_double(p.E), C.c_double(p.nu),
                                C.c_double(p.thetaC), C.c_double(p.thetaS))
        if scene.mesh is not None:
            m = scene.mesh; self.nv, self.nf = m.nv, m.nf
            vB = np.concatenate([colmajor(m.vB[:, a, :]).ravel() for a in range(3)])
            eB = np.concatenate([colmajor(m.eB[:, a, :]).ravel() for a in range(3)])
            ed = np.concatenate([colmajor(m.ed[a]).ravel() for a in range(3)])
            eD = np.concatenate([colmajor(m.eD[a]).ravel() for a in range(3)])
            faces = np.ascontiguousarray(m.faces.T.astype(np.int32))
            fixed = None if m.fixed is None else np.ascontiguousarray(m.fixed, np.float64)
            L.orc_set_mesh(self.h, C.c_long(m.nv), C.c_long(m.nf), _p(colmajor(m.vx)), _p(colmajor(m.vv)),
                           _p(np.ascontiguousarray(m.vm)), _p(np.ascontiguousarray(m.vvol)), _p(vB),
                           faces.ctypes.data_as(C.POINTER(C.c_int)), _p(colmajor(m.ev)), _p(np.ascontiguousarray(m.em)),
                           _p(np.ascontiguousarray(m.evol)), _p(eB), _p(ed), _p(eD), _p(fixed),
                           C.c_double(m.mu), C.c_double(m.lam), C.c_double(m.shear), C.c_double(m.stiff), C.c_double(m.fric))
        self._set_levelset(scene)

    def _set_levelset(self, scene):
        if scene.levelset.kind != 0:
            par = np.ascontiguousarray(scene.levelset.params, np.float64)
            self.L.orc_set_levelset(self.h, C.c_int(scene.levelset.kind), _p(par))

    def __del__(self):
        try:
            self.L.orc_destroy(self.h)
        except Exception:
            pass

    # ---- stepping
    def init(self): self.L.orc_init(self.h)
    def substep(self): return self.L.orc_substep(self.h)
    def rebuild_weights(self): self.L.orc_rebuild_weights(self.h)
    def p2g(self, first=False): self.L.orc_p2g(self.h, C.c_int(int(first)))
    def stage_forces(self, dt): self.L.orc_stage_forces(self.h, C.c_double(dt))
    def stage_grid_update(self, dt): self.L.orc_stage_grid_update(self.h, C.c_double(dt))
    def cfl_condition(self): return self.L.orc_cfl_condition(self.h)
    def stage_collide(self): self.L.orc_stage_collide(self.h)
    def stage_g2p(self, dt): self.L.orc_stage_g2p(self.h, C.c_double(dt))
    def set_dt(self, dt): self.L.orc_set_dt(self.h, C.c_double(dt))
    @property
    def dt(self): return self.L.orc_get_dt(self.h)
    @property
    def time(self): return self.L.orc_get_time(self.h)
    @property
    def frame(self): return self.L.orc_get_frame(self.h)

    def timers(self):
        out = np.zeros(5); self.L.orc_get_timers(self.h, _p(out))
        return dict(zip(("forces", "grid", "g2p", "weights", "p2g"), out))

    # ---- state
    def particles(self):
        n = self.np
        b = {k: np.empty(3 * n) for k in ("x", "v", "B1", "B2", "B3")}
        FE = np.empty(9 * n); FP = np.empty(9 * n); vol = np.empty(n); q = np.empty(n)
        self.L.orc_get_particles(self.h, _p(b["x"]), _p(b["v"]), _p(b["B1"]), _p(b["B2"]), _p(b["B3"]), _p(FE), _p(FP), _p(vol), _p(q))
        B = np.stack([from_colmajor(b["B1"], n), from_colmajor(b["B2"], n), from_colmajor(b["B3"], n)], axis=1)
        return dict(x=from_colmajor(b["x"], n), v=from_colmajor(b["v"], n), B=B, FE=mats_from_colmajor(FE, n),
                    FP=mats_from_colmajor(FP, n), vol=vol, q=q)

    def grid(self):
        ng = self.ng
        m = np.empty(ng); v = np.empty(3 * ng); f = np.empty(3 * ng); vt = np.empty(3 * ng)
        self.L.orc_get_grid(self.h, _p(m), _p(v), _p(f), _p(vt))
        return dict(m=m, v=from_colmajor(v, ng), f=from_colmajor(f, ng), vt=from_colmajor(vt, ng))

    def set_grid(self, m=None, v=None, f=None):
        cm = lambda a: None if a is None else colmajor(a)
        mm = None if m is None else np.ascontiguousarray(m, np.float64); vv = cm(v); ff = cm(f)
        self.L.orc_set_grid(self.h, _p(mm), _p(vv), _p(ff))

    def compute_volumes(self): self.L.orc_compute_volumes(self.h)

    def set_particles(self, p):
        """Replace the particle set (scenes.Particles)."""
        self.np = p.n
        arrs = [colmajor(p.x), colmajor(p.v), colmajor(p.B[:, 0, :]), colmajor(p.B[:, 1, :]), colmajor(p.B[:, 2, :]),
                mats_colmajor(p.FE), mats_colmajor(p.FP), np.ascontiguousarray(p.m, np.float64),
                np.ascontiguousarray(p.vol, np.float64), np.ascontiguousarray(p.q, np.float64)]
        self.L.orc_set_particles(self.h, C.c_long(p.n), *[_p(a) for a in arrs], C.c_double(p.E), C.c_double(p.nu),
                                 C.c_double(p.thetaC), C.c_double(p.thetaS))

    def mesh(self):
        nv, nf = self.nv, self.nf
        vx = np.empty(3 * nv); vv = np.empty(3 * nv); vB = np.empty(9 * nv)
        ex = np.empty(3 * nf); ev = np.empty(3 * nf); eB = np.empty(9 * nf); ed = np.empty(9 * nf)
        self.L.orc_get_mesh(self.h, _p(vx), _p(vv), _p(vB), _p(ex), _p(ev), _p(eB), _p(ed))
        un = lambda buf, n: np.stack([from_colmajor(buf[3 * n * a:3 * n * (a + 1)], n) for a in range(3)], axis=0)
        return dict(vx=from_colmajor(vx, nv), vv=from_colmajor(vv, nv), vB=un(vB, nv).transpose(1, 0, 2),
                    ex=from_colmajor(ex, nf), ev=from_colmajor(ev, nf), eB=un(eB, nf).transpose(1, 0, 2), ed=un(ed, nf))


# ---- math hooks for unit tests
def svd3(F):
    L = lib(); Fc = np.ascontiguousarray(np.asarray(F, np.float64).T).ravel()
    U = np.empty(9); s = np.empty(3); V = np.empty(9)
    L.orc_svd3(_p(Fc), _p(U), _p(s), _p(V))
    return U.reshape(3, 3).T.copy(), s, V.reshape(3, 3).T.copy()


def svd2(A):
    L = lib(); Ac = np.ascontiguousarray(np.asarray(A, np.float64).T).ravel()
    U = np.empty(4); s = np.empty(2); V = np.empty(4)
    L.orc_svd2(_p(Ac), _p(U), _p(s), _p(V))
    return U.reshape(2, 2).T.copy(), s, V.reshape(2, 2).T.copy()


def gram_schmidt(A):
    L = lib(); Ac = np.ascontiguousarray(np.asarray(A, np.float64).T).ravel()
    Q = np.empty(9); R = np.empty(9)
    L.orc_gram_schmidt(_p(Ac), _p(Q), _p(R))
    return Q.reshape(3, 3).T.copy(), R.reshape(3, 3).T.copy()


def cubic_bspline(x): return lib().orc_cubic_bspline(C.c_double(x))
def dcubic_bspline(x): return lib().orc_dcubic_bspline(C.c_double(x))
