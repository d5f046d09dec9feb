"""TEST-ONLY backend for anisotropicelastoplasticity_b200.distributed: one slab of the domain on the CPU oracle, with the
same interface as GpuSlabBackend (torch CPU tensors as communication buffers).  Lets the real multi-rank driver
(SlabSolver over torch.distributed/gloo, or LocalSlabGroup) run without a GPU and be compared with the single-process oracle."""
import copy

import numpy as np
import torch

from anisotropicelastoplasticity_b200 import scenes as sc
from oracle.oracle_py import Oracle

REC = 37   # id, x3, v3, B9, FE9, FP9, m, vol, q


class OracleSlabBackend:
    def __init__(self, scene, plan, rank):
        self.scene = scene; self.plan = plan; self.rank = rank
        self.axis, self.lo, self.hi = plan.slab(rank)
        g = scene.grid; self.g = g; self.res = [int(r) for r in g.res]; self.h = g.h; self.hmin = float(g.h.min())
        p = scene.particles
        cells = np.floor((p.x[:, self.axis] - g.mn[self.axis]) / g.h[self.axis]).astype(np.int64)
        idx = np.nonzero((cells >= self.lo) & (cells < self.hi))[0]
        self.ids = idx.astype(np.int64)
        self.P = sc.Particles(x=p.x[idx].copy(), v=p.v[idx].copy(), B=p.B[idx].copy(), FE=p.FE[idx].copy(), FP=p.FP[idx].copy(), m=p.m[idx].copy(),
                              vol=p.vol[idx].copy(), q=p.q[idx].copy(), E=p.E, nu=p.nu, thetaC=p.thetaC, thetaS=p.thetaS)
        shell = copy.copy(scene); shell.particles = self.P
        self.o = Oracle(shell)
        self.cfl = scene.cfl; self.rate = 3e2; self.frame_dt = 1.0 / 60.0
        self.dt = 0.0; self.t = 0.0; self.inner_t = 0.0; self.frame = 0; self.vmax = 0.0
        self.has = {0: self.lo > 0, 1: self.hi < self.res[self.axis]}

    # ---- plane helpers: grid arrays as [k][j][i]
    def _planes(self, side):
        b = self.lo if side == 0 else self.hi
        return [p for p in (b - 1, b, b + 1)]

    def _take(self, arr3, side):
        out = []
        for pl in self._planes(side):
            if 0 <= pl < self.res[self.axis]:
                out.append(np.take(arr3, pl, axis=2 - self.axis))
            else:
                out.append(np.zeros_like(np.take(arr3, 0, axis=2 - self.axis)))
        return np.stack(out, axis=0)

    def _put_add(self, arr3, side, planes):
        for n, pl in enumerate(self._planes(side)):
            if 0 <= pl < self.res[self.axis]:
                sl = [slice(None)] * arr3.ndim; sl[2 - self.axis] = pl
                arr3[tuple(sl)] += planes[n]

    def _grid4(self, what):
        g = self.o.grid(); nz, ny, nx = self.res[2], self.res[1], self.res[0]
        if what == 0:
            a = np.concatenate([g["m"][:, None], g["m"][:, None] * g["v"]], axis=1)
        else:
            f = g["f"].copy(); f[:, 2] += 9.8 * g["m"]           # strip the gravity term the oracle folds into f (HS:457)
            a = np.concatenate([f, np.zeros((f.shape[0], 1))], axis=1)
        return a.reshape(nz, ny, nx, 4), g

    # ---- backend interface
    def init_begin(self): self.o.rebuild_weights(); self.o.p2g(False)
    def init_volumes(self): self.o.compute_volumes(); self.vmax = self.o.cfl_condition() * self.hmin
    def init_dt(self): self.dt = self.cfl / max(self.rate, self.vmax / self.hmin)

    def halo_pack(self, what, side):
        a, _ = self._grid4(what)
        return torch.from_numpy(np.ascontiguousarray(self._take(a, side)))

    def halo_recv_buffer(self, what, side):
        a, _ = self._grid4(what)
        return torch.zeros(self._take(a, side).shape, dtype=torch.float64)

    def halo_add(self, what, side, t):
        a, g = self._grid4(what)
        a = a.copy(); self._put_add(a, side, t.numpy())
        flat = a.reshape(-1, 4)
        if what == 0:
            m = flat[:, 0]; v = np.zeros_like(flat[:, 1:]); nz = m > 0.0
            v[nz] = flat[nz, 1:] / m[nz, None]
            self.o.set_grid(m=m, v=v)
        else:
            f = flat[:, :3].copy(); f[:, 2] -= 9.8 * g["m"]
            self.o.set_grid(f=f)

    def vmax_get(self): return torch.tensor([self.vmax], dtype=torch.float64)
    def vmax_set(self, t): self.vmax = float(t.item())

    def step_forces(self): self.o.stage_forces(self.dt)

    def step_grid(self):
        self.o.stage_grid_update(self.dt); self.vmax = self.o.cfl_condition() * self.hmin; self.o.stage_collide()

    def step_g2p(self):
        dt = self.cfl / max(self.rate, self.vmax / self.hmin)             # HS:878-892
        if self.inner_t + dt >= self.frame_dt:
            dt = self.frame_dt - self.inner_t; self.t += self.frame_dt; self.inner_t = 0.0; self.frame += 1
        else:
            self.inner_t += dt
        self.dt = dt
        self.o.stage_g2p(dt)

    def _pull(self):
        p = self.o.particles()
        self.P.x, self.P.v, self.P.B, self.P.FE, self.P.FP, self.P.vol, self.P.q = p["x"], p["v"], p["B"], p["FE"], p["FP"], p["vol"], p["q"]

    def migrate_extract(self):
        self._pull(); P = self.P
        cells = np.floor((P.x[:, self.axis] - self.g.mn[self.axis]) / self.g.h[self.axis]).astype(np.int64)
        out = []
        keep = np.ones(P.n, bool)
        for side in (0, 1):
            sel = (cells < self.lo) if side == 0 else (cells >= self.hi)
            n = int(sel.sum())
            rec = np.concatenate([self.ids[sel, None].astype(np.float64), P.x[sel], P.v[sel], P.B[sel].reshape(n, 9), P.FE[sel].reshape(n, 9),
                                  P.FP[sel].reshape(n, 9), P.m[sel, None], P.vol[sel, None], P.q[sel, None]], axis=1)
            out.append(torch.from_numpy(np.ascontiguousarray(rec))); keep &= ~sel
        for k in ("x", "v", "B", "FE", "FP", "m", "vol", "q"):
            setattr(P, k, getattr(P, k)[keep])
        self.ids = self.ids[keep]
        return out[0], out[1]

    # overlapped-migration interface of SlabSolver (the GPU backend hides the count round trip behind the resident P2G; here the
    # order of the same operations is simply kept)
    def migrate_begin(self):
        self._mig_out = self.migrate_extract()
        self._c_out = torch.tensor([self._mig_out[0].shape[0], self._mig_out[1].shape[0]], dtype=torch.int64)
        self._c_in = torch.zeros(2, dtype=torch.int64)
        self._defer_p2g = True
        return self._c_out, self._c_in

    def migrate_counts_to_host(self): return None

    def migrate_end(self, ev):
        return self._mig_out, (int(self._c_in[0]), int(self._c_in[1]))

    def step_p2g_arrivals(self, count):
        self._defer_p2g = False
        self.step_p2g()

    def migrate_recv_buffer(self, side, n): return torch.zeros((n, REC), dtype=torch.float64)

    def migrate_insert(self, a, b):
        P = self.P
        for t in (a, b):
            if t.shape[0] == 0:
                continue
            r = t.numpy(); n = r.shape[0]
            self.ids = np.concatenate([self.ids, r[:, 0].astype(np.int64)])
            P.x = np.concatenate([P.x, r[:, 1:4]]); P.v = np.concatenate([P.v, r[:, 4:7]])
            P.B = np.concatenate([P.B, r[:, 7:16].reshape(n, 3, 3)]); P.FE = np.concatenate([P.FE, r[:, 16:25].reshape(n, 3, 3)])
            P.FP = np.concatenate([P.FP, r[:, 25:34].reshape(n, 3, 3)]); P.m = np.concatenate([P.m, r[:, 34]])
            P.vol = np.concatenate([P.vol, r[:, 35]]); P.q = np.concatenate([P.q, r[:, 36]])

    def step_p2g(self):
        if getattr(self, "_defer_p2g", False):
            return                                               # the whole P2G happens after the arrivals are appended
        self.o.set_particles(self.P); self.o.rebuild_weights(); self.o.p2g(False)

    def sync(self): pass

    def particles_local(self):
        self._pull(); P = self.P
        return dict(ids=self.ids.copy(), x=P.x.copy(), v=P.v.copy(), B=P.B.copy(), FE=P.FE.copy(), FP=P.FP.copy(), vol=P.vol.copy(), q=P.q.copy())
