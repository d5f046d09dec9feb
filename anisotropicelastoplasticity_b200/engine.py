"""Python host mirror of the reference's solver surface on top of the C ABI (include/aep_b200.h).

`Engine` plays the role of HybridSolver (HybridSolver.h:27-95) for one GPU: it is handed a Scene
(ParticleSystem + RegularGrid + optional LagrangianMesh + level set), uploads it, and steps it on the device.
Method names follow the reference's private stages so the parity tests read like its call sequence
(HybridSolver.cpp:867-1032).  Everything numerical happens inside libaep_b200.so.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .scenes import Scene, colmajor, from_colmajor, mats_colmajor, mats_from_colmajor


def _p(a):
    return None if a is None else a.ctypes.data_as(capi.dp)


class Engine:
    def __init__(self, scene: Scene, device: int = 0, particle_capacity: int = 0, slab=None, dt_rate_floor=None, sort_every=None, sort_bricks=None, scatter_strips=None,
                 vmax_min_mass_fraction=None, coulomb_friction=None, use_graph=None, sort_cost_threshold=None):
        self.L = capi.load()
        cfg = capi.Config(); capi.check(self.L.aep_default_config(C.byref(cfg)))
        g = scene.grid
        cfg.device = device; cfg.material = scene.material; cfg.cfl = scene.cfl
        for a in range(3):
            cfg.grid_min[a] = float(g.mn[a]); cfg.grid_max[a] = float(g.mx[a]); cfg.res[a] = int(g.res[a])
        cfg.particle_capacity = particle_capacity
        if dt_rate_floor is not None:
            cfg.dt_rate_floor = float(dt_rate_floor)
        if sort_every is not None:
            cfg.sort_every = int(sort_every)
        if scatter_strips is not None:
            cfg.scatter_strips = int(scatter_strips)
        if sort_bricks is not None:
            cfg.sort_bricks = int(sort_bricks)
        if vmax_min_mass_fraction is not None:
            cfg.vmax_min_mass_fraction = float(vmax_min_mass_fraction)      # opt-in, NOT the reference's dt rule
        if coulomb_friction is not None:
            cfg.coulomb_friction = int(coulomb_friction)                    # opt-in, NOT the reference's collider
        if use_graph is not None:
            cfg.use_graph = int(use_graph)
        if sort_cost_threshold is not None:
            cfg.sort_cost_threshold = float(sort_cost_threshold)
        if slab is not None:
            cfg.slab_axis, cfg.slab_lo, cfg.slab_hi = slab
        self.cfg = cfg
        self.h = C.c_void_p()
        capi.check(self.L.aep_create(C.byref(self.h), C.byref(cfg)))
        self.ng = g.n_nodes; self.nv = 0; self.nf = 0
        self.scene_name = scene.name
        if scene.particles is not None:
            self.upload_particles(scene.particles)
        if scene.mesh is not None:
            self.upload_mesh(scene.mesh)
        if scene.levelset.kind != 0:
            par = np.ascontiguousarray(scene.levelset.params, np.float64)
            capi.check(self.L.aep_set_levelset_analytic(self.h, int(scene.levelset.kind), _p(par)), self.h)

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.L.aep_destroy(self.h); self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ upload
    @staticmethod
    def pack_particles(p):
        """Host arrays in the reference's memory layouts (what ParticleSystem's Eigen members hold)."""
        return [colmajor(p.x), colmajor(p.v), colmajor(p.B[:, 0, :]), colmajor(p.B[:, 1, :]), colmajor(p.B[:, 2, :]),
                mats_colmajor(p.FE), mats_colmajor(p.FP), np.ascontiguousarray(p.m, np.float64),
                np.ascontiguousarray(p.vol, np.float64), np.ascontiguousarray(p.q, np.float64)]

    def upload_particles(self, p, packed=None):
        arrs = packed if packed is not None else self.pack_particles(p)
        capi.check(self.L.aep_upload_particles(self.h, p.n, *[_p(a) for a in arrs], p.E, p.nu, p.thetaC, p.thetaS), self.h)

    def upload_packed(self, n, arrs, E, nu, thetaC, thetaS):
        capi.check(self.L.aep_upload_particles(self.h, n, *[_p(a) for a in arrs], E, nu, thetaC, thetaS), self.h)

    def upload_mesh(self, m):
        self.nv, self.nf = m.nv, m.nf
        vB = np.concatenate([colmajor(m.vB[:, a, :]).ravel() for a in range(3)])
        eB = np.concatenate([colmajor(m.eB[:, a, :]).ravel() for a in range(3)])
        ed = np.concatenate([colmajor(m.ed[a]).ravel() for a in range(3)])
        eD = np.concatenate([colmajor(m.eD[a]).ravel() for a in range(3)])
        faces = np.ascontiguousarray(m.faces.T.astype(np.int32))
        fixed = None if m.fixed is None else np.ascontiguousarray(m.fixed, np.float64)
        capi.check(self.L.aep_upload_mesh(self.h, m.nv, m.nf, _p(colmajor(m.vx)), _p(colmajor(m.vv)),
                                          _p(np.ascontiguousarray(m.vm, np.float64)), _p(np.ascontiguousarray(m.vvol, np.float64)), _p(vB),
                                          faces.ctypes.data_as(C.POINTER(C.c_int32)), _p(colmajor(m.ev)),
                                          _p(np.ascontiguousarray(m.em, np.float64)), _p(np.ascontiguousarray(m.evol, np.float64)), _p(eB), _p(ed), _p(eD),
                                          _p(fixed), m.mu, m.lam, m.shear, m.stiff, m.fric), self.h)

    def set_collider_motion(self, velocity):
        v = None if velocity is None else np.ascontiguousarray(velocity, np.float64)
        capi.check(self.L.aep_set_collider_motion(self.h, _p(v)), self.h)

    def counters(self):
        a = [C.c_int64() for _ in range(4)]
        capi.check(self.L.aep_get_counters(self.h, *[C.byref(x) for x in a]), self.h)
        return dict(sorts=a[0].value, slots=a[1].value, dead=a[2].value, moved_since_sort=a[3].value)

    def migration(self):
        a = [C.c_int64() for _ in range(2)]
        capi.check(self.L.aep_get_migration(self.h, *[C.byref(x) for x in a]), self.h)
        return dict(sent=a[0].value, received=a[1].value)

    def set_levelset_samples(self, inside, normal):
        inside = np.ascontiguousarray(inside, np.uint8); normal = colmajor(normal)
        capi.check(self.L.aep_set_levelset_samples(self.h, inside.ctypes.data_as(C.POINTER(C.c_uint8)), _p(normal)), self.h)

    # ------------------------------------------------------------------ stepping
    def init(self): capi.check(self.L.aep_init(self.h), self.h)
    def substep(self): capi.check(self.L.aep_substep(self.h), self.h)
    def run(self, n): capi.check(self.L.aep_run(self.h, int(n)), self.h)
    def sync(self): capi.check(self.L.aep_sync(self.h), self.h)

    def run_frames(self, n_frames, max_substeps=1 << 30):
        done = C.c_int64(0)
        capi.check(self.L.aep_run_frames(self.h, int(n_frames), int(max_substeps), C.byref(done)), self.h)
        return done.value

    def p2g(self, first=False): capi.check(self.L.aep_p2g(self.h, int(first)), self.h)
    def stage_forces(self, dt): capi.check(self.L.aep_stage_forces(self.h, float(dt)), self.h)
    def stage_grid(self, dt): capi.check(self.L.aep_stage_grid(self.h, float(dt)), self.h)
    def stage_g2p(self, dt): capi.check(self.L.aep_stage_g2p(self.h, float(dt)), self.h)
    def set_dt(self, dt): capi.check(self.L.aep_set_dt(self.h, float(dt)), self.h)
    def set_fixed_dt(self, dt): capi.check(self.L.aep_set_fixed_dt(self.h, float(dt)), self.h)

    # ------------------------------------------------------------------ checkpoint / restart
    def checkpoint(self):
        """Everything needed to continue this run elsewhere: particle and mesh state in the reference's layouts + the clock."""
        c = self.clock()
        out = dict(clock=np.array([c["dt"], c["t"], c["inner_t"], c["frame"], c["substeps"], c["escaped"]], np.float64))
        if self.n_particles:
            out.update({"p_" + k: v for k, v in self.particles().items()})
        if self.nv:
            out.update({"m_" + k: v for k, v in self.mesh().items()})
        return out

    @classmethod
    def resume(cls, scene: Scene, ckpt, **kw):
        """A new context that continues from `ckpt` (Engine.checkpoint(), or np.load of its np.savez).  `scene` supplies what a
        checkpoint does not carry: grid, material constants, masses, mesh topology and rest state, collider."""
        import copy
        s = copy.deepcopy(scene)
        if s.particles is not None:
            p = s.particles
            p.x, p.v, p.B, p.FE, p.FP, p.vol, p.q = (np.array(ckpt["p_" + k]) for k in ("x", "v", "B", "FE", "FP", "vol", "q"))
        if s.mesh is not None:
            m = s.mesh
            m.vx, m.vv, m.vB, m.ev, m.eB, m.ed = (np.array(ckpt["m_" + k]) for k in ("vx", "vv", "vB", "ev", "eB", "ed"))
        e = cls(s, **kw)
        capi.check(e.L.aep_resume(e.h), e.h)
        ck = [float(v) for v in np.asarray(ckpt["clock"])]
        dt, t, inner_t, frame, substeps = ck[:5]
        capi.check(e.L.aep_set_clock(e.h, dt, t, inner_t, int(frame), int(substeps)), e.h)
        if len(ck) > 5 and ck[5] > 0:
            capi.check(e.L.aep_set_escaped(e.h, int(ck[5])), e.h)
        return e

    def clock(self):
        dt = C.c_double(); t = C.c_double(); it = C.c_double(); fr = C.c_int32(); ss = C.c_int64(); vm = C.c_double(); esc = C.c_int64()
        capi.check(self.L.aep_get_clock(self.h, C.byref(dt), C.byref(t), C.byref(it), C.byref(fr), C.byref(ss), C.byref(vm), C.byref(esc)), self.h)
        return dict(dt=dt.value, t=t.value, inner_t=it.value, frame=fr.value, substeps=ss.value, vmax=vm.value, escaped=esc.value)

    @property
    def dt(self): return self.clock()["dt"]

    @property
    def n_particles(self): return int(self.L.aep_num_particles(self.h))

    @property
    def kernel_launches(self): return int(self.L.aep_kernel_launches(self.h))

    @property
    def stream(self): return self.L.aep_stream(self.h)

    def profile(self, enable=True): capi.check(self.L.aep_profile(self.h, int(enable)), self.h)

    def timers(self):
        ms = np.zeros(capi.NUM_STAGES); calls = np.zeros(capi.NUM_STAGES, np.int64)
        capi.check(self.L.aep_get_timers(self.h, _p(ms), calls.ctypes.data_as(capi.i64p)), self.h)
        return {s: (ms[i], int(calls[i])) for i, s in enumerate(capi.STAGES) if s != "_"}

    # ------------------------------------------------------------------ download
    def particles(self):
        n = self.n_particles
        b = {k: np.empty(3 * n) for k in ("x", "v", "B1", "B2", "B3")}
        FE = np.empty(9 * n); FP = np.empty(9 * n); vol = np.empty(n); q = np.empty(n)
        capi.check(self.L.aep_download_particles(self.h, _p(b["x"]), _p(b["v"]), _p(b["B1"]), _p(b["B2"]), _p(b["B3"]),
                                                 _p(FE), _p(FP), _p(vol), _p(q)), self.h)
        B = np.stack([from_colmajor(b["B1"], n), from_colmajor(b["B2"], n), from_colmajor(b["B3"], n)], axis=1)
        return dict(x=from_colmajor(b["x"], n), v=from_colmajor(b["v"], n), B=B, FE=mats_from_colmajor(FE, n),
                    FP=mats_from_colmajor(FP, n), vol=vol, q=q)

    def positions_f32(self):
        out = np.empty((self.n_particles, 3), np.float32)
        capi.check(self.L.aep_download_positions_f32(self.h, out.ctypes.data_as(C.POINTER(C.c_float))), self.h)
        return out

    def grid(self):
        ng = self.ng
        m = np.empty(ng); v = np.empty(3 * ng); f = np.empty(3 * ng); vt = np.empty(3 * ng)
        capi.check(self.L.aep_download_grid(self.h, _p(m), _p(v), _p(f), _p(vt)), self.h)
        return dict(m=m, v=from_colmajor(v, ng), f=from_colmajor(f, ng), vt=from_colmajor(vt, ng))

    def mesh(self):
        nv, nf = self.nv, self.nf
        vx = np.empty(3 * nv); vv = np.empty(3 * nv); vB = np.empty(9 * nv)
        ex = np.empty(3 * nf); ev = np.empty(3 * nf); eB = np.empty(9 * nf); ed = np.empty(9 * nf)
        capi.check(self.L.aep_download_mesh(self.h, _p(vx), _p(vv), _p(vB), _p(ex), _p(ev), _p(eB), _p(ed)), self.h)
        un = lambda buf, n: np.stack([from_colmajor(buf[3 * n * a:3 * n * (a + 1)], n) for a in range(3)], axis=0)
        return dict(vx=from_colmajor(vx, nv), vv=from_colmajor(vv, nv), vB=un(vB, nv).transpose(1, 0, 2),
                    ex=from_colmajor(ex, nf), ev=from_colmajor(ev, nf), eB=un(eB, nf).transpose(1, 0, 2), ed=un(ed, nf))

    def grid_activity(self):
        b = C.c_int64(); n = C.c_int64()
        capi.check(self.L.aep_grid_activity(self.h, C.byref(b), C.byref(n)), self.h)
        return b.value, n.value

    def stats(self):
        com = np.zeros(3); ke = C.c_double(); jp = C.c_double(); mass = C.c_double()
        capi.check(self.L.aep_stats(self.h, _p(com), C.byref(ke), C.byref(jp), C.byref(mass)), self.h)
        return dict(com=com, ke=ke.value, jp=jp.value, mass=mass.value)
