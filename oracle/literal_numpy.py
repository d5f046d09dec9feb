"""oracle/literal_numpy.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Independent float64 transcription of the reference's *sparse-matrix formulation*
(AnisotropicElastoplasticity/HybridSolver.cpp) on numpy / scipy.sparse / numpy.linalg.svd.
scipy.sparse maps 1:1 onto the Eigen::SparseMatrix algebra the reference uses, and LAPACK's SVD has the
same contract as Eigen::JacobiSVD (s >= 0, sorted descending), so this file pins the C++ restatement
(oracle/mpm_oracle.cpp, direct-stencil form) through a second, structurally different implementation.
It also generates the golden vectors under tests/golden/ (oracle/make_golden.py).

Only tests/ and oracle/make_golden.py import this module.  Slow: small scenes only.

Citations: HS = HybridSolver.cpp, LM = LagrangianMesh.cpp, RG = RegularGrid.cpp, IP = interpolation.cpp,
GE = geometry.cpp, LS = LevelSet.cpp (all under /root/reference/AnisotropicElastoplasticity/).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

SNOW, SAND = 0, 1


def cubic_B_spline(x):
    """IP:9-16."""
    ax = np.abs(x)
    return np.where(ax >= 2.0, 0.0,
                    np.where(ax >= 1.0, -1.0 / 6.0 * ax * ax * ax + ax * ax - 2.0 * ax + 4.0 / 3.0,
                             0.5 * ax * ax * ax - ax * ax + 2.0 / 3.0))


def Dcubic_B_spline(x):
    """IP:18-33."""
    return np.where(x >= 2.0, 0.0,
           np.where(x >= 1.0, -0.5 * x * x + 2.0 * x - 2.0,
           np.where(x >= 0.0, 1.5 * x * x - 2.0 * x,
           np.where(x >= -1.0, -1.5 * x * x - 2.0 * x,
           np.where(x >= -2.0, 0.5 * x * x + 2.0 * x + 2.0, 0.0)))))


def gram_schmidt(A):
    """GE:31-62 on a batch A[n,3,3] (columns d1,d2,d3). Returns Q, R."""
    d1, d2, d3 = A[:, :, 0], A[:, :, 1], A[:, :, 2]
    r11 = np.linalg.norm(d1, axis=1); q1 = d1 / r11[:, None]
    r12 = np.einsum('ni,ni->n', d2, q1)
    q2 = d2 - r12[:, None] * q1
    r22 = np.linalg.norm(q2, axis=1); q2 = q2 / r22[:, None]
    r13 = np.einsum('ni,ni->n', d3, q1); r23 = np.einsum('ni,ni->n', d3, q2)
    q3 = d3 - r13[:, None] * q1 - r23[:, None] * q2
    r33 = np.linalg.norm(q3, axis=1); q3 = q3 / r33[:, None]
    Q = np.stack([q1, q2, q3], axis=2)
    R = np.zeros_like(A)
    R[:, 0, 0], R[:, 0, 1], R[:, 0, 2], R[:, 1, 1], R[:, 1, 2], R[:, 2, 2] = r11, r12, r13, r22, r23, r33
    return Q, R


class Grid:
    """RG:117-177."""

    def __init__(self, mn, mx, res):
        self.mn = np.asarray(mn, float); self.mx = np.asarray(mx, float); self.res = np.asarray(res, int)
        self.h = (self.mx - self.mn) / self.res
        self.Ng = int(np.prod(self.res))
        k, j, i = np.meshgrid(np.arange(res[2]), np.arange(res[1]), np.arange(res[0]), indexing='ij')
        self.positions = np.stack([self.mn[0] + i.ravel() * self.h[0], self.mn[1] + j.ravel() * self.h[1],
                                   self.mn[2] + k.ravel() * self.h[2]], axis=1)      # index = k*nx*ny + j*nx + i
        self.masses = np.zeros(self.Ng); self.velocities = np.zeros((self.Ng, 3)); self.forces = np.zeros((self.Ng, 3))

    def cfl_condition(self):
        return np.linalg.norm(self.velocities, axis=1).max() / self.h.min()           # RG:188-200, RegularGrid.h:60


class LiteralSolver:
    def __init__(self, grid: Grid, material=SAND, cfl=0.3):
        self.rg = grid; self.material = material; self.cfl = cfl
        self.ps = None; self.mesh = None
        self.phi = None; self.dphi = None            # callables on (n,3) arrays -> (n,), (n,3)
        self.t = 0.0; self.inner_t = 0.0; self.dt = 0.0; self.frame_no = 0

    # ------------------------------------------------------------------ HS:18-97
    def weights(self, pos):
        rg = self.rg; n = pos.shape[0]
        pr = pos - rg.mn
        fl = (pr / rg.h).astype(np.int64)             # static_cast<int>: truncation toward zero
        offs = np.arange(-2, 3)
        rows, cols, w, d1, d2, d3 = [], [], [], [], [], []
        for oi in offs:
            for oj in offs:
                for ok in offs:
                    i = fl[:, 0] + oi; j = fl[:, 1] + oj; k = fl[:, 2] + ok
                    ing = (0 <= i) & (i < rg.res[0]) & (0 <= j) & (j < rg.res[1]) & (0 <= k) & (k < rg.res[2])
                    di = cubic_B_spline(pr[:, 0] / rg.h[0] - i); dj = cubic_B_spline(pr[:, 1] / rg.h[1] - j)
                    dk = cubic_B_spline(pr[:, 2] / rg.h[2] - k)
                    ddi = Dcubic_B_spline(pr[:, 0] / rg.h[0] - i) / rg.h[0]
                    ddj = Dcubic_B_spline(pr[:, 1] / rg.h[1] - j) / rg.h[1]
                    ddk = Dcubic_B_spline(pr[:, 2] / rg.h[2] - k) / rg.h[2]
                    keep = ing & (di > 0) & (dj > 0) & (dk > 0)
                    idx = k * rg.res[0] * rg.res[1] + j * rg.res[0] + i
                    p = np.nonzero(keep)[0]
                    rows.append(p); cols.append(idx[p])
                    w.append((di * dj * dk)[p]); d1.append((ddi * dj * dk)[p]); d2.append((di * ddj * dk)[p]); d3.append((di * dj * ddk)[p])
        rows = np.concatenate(rows); cols = np.concatenate(cols)
        mk = lambda v: sp.csc_matrix((np.concatenate(v), (rows, cols)), shape=(n, rg.Ng))   # setFromTriplets
        return mk(w), mk(d1), mk(d2), mk(d3)

    def rebuild_weights(self):
        if self.ps is not None:
            self.om, self.dom1, self.dom2, self.dom3 = self.weights(self.ps['x'])
        if self.mesh is not None:
            self.vom, self.dvom1, self.dvom2, self.dvom3 = self.weights(self.mesh['vx'])
            self.eom, self.deom1, self.deom2, self.deom3 = self.weights(self.mesh['ex'])

    # ------------------------------------------------------------------ HS:113-250
    def particle_to_grid(self, first):
        rg = self.rg
        rg.masses = np.zeros(rg.Ng)
        mom = np.zeros((rg.Ng, 3))
        sets = []
        if self.ps is not None:
            p = self.ps; sets.append((self.om, p['m'], p['v'], p['B1'], p['B2'], p['B3'], p['x']))
        if self.mesh is not None:
            m = self.mesh
            sets.append((self.vom, m['vm'], m['vv'], m['vB1'], m['vB2'], m['vB3'], m['vx']))
            sets.append((self.eom, m['em'], m['ev'], m['eB1'], m['eB2'], m['eB3'], m['ex']))
        for om, mass, vel, *_ in sets:
            rg.masses = rg.masses + om.T @ mass                                           # HS:118-124
        for om, mass, vel, *_ in sets:
            mom = mom + om.T @ (sp.diags(mass) @ vel)                                     # HS:134-140
        hmin = rg.h.min(); ratio = 3.0 / hmin / hmin                                      # HS:175-177
        for om, mass, vel, b1, b2, b3, pos in sets:                                       # HS:178-203 (optimised algebra)
            for a, Ba in enumerate((b1, b2, b3)):
                cell = ratio * (om.T @ (sp.diags(mass) @ Ba))
                mom[:, a] += (cell * rg.positions).sum(axis=1)
                mom[:, a] -= ratio * (om.T @ (mass * (Ba * pos).sum(axis=1)))
        rg.velocities = np.zeros((rg.Ng, 3))
        nz = rg.masses > 0.0
        rg.velocities[nz] = mom[nz] / rg.masses[nz, None]                                 # HS:233-240
        if first and self.ps is not None:                                                 # HS:242-249
            p = self.ps
            p['dens'] = self.om @ rg.masses / np.prod(rg.h)
            p['vol'] = p['m'] * (1.0 / p['dens'])

    # ------------------------------------------------------------------ LM:382-460
    def cloth_in_plane(self):
        m = self.mesh; nf = m['faces'].shape[0]
        Dm = np.stack([m['eD1'], m['eD2'], m['eD3']], axis=2); dm = np.stack([m['ed1'], m['ed2'], m['ed3']], axis=2)
        Q, R = gram_schmidt(dm); Q0, R0 = gram_schmidt(Dm)
        vf = np.zeros_like(m['vx']); pk = np.zeros((nf, 2, 2))
        for f in range(nf):
            inR = R[f, :2, :2]; r0 = R0[f, :2, :2]
            inv = np.array([[1.0 / r0[0, 0], -r0[0, 1] / r0[0, 0] / r0[1, 1]], [0.0, 1.0 / r0[1, 1]]])   # GE:67-73
            ref = inv @ inR
            invRefMulDet = np.array([[ref[1, 1], -ref[0, 1]], [0.0, ref[0, 0]]])
            U, s, Vt = np.linalg.svd(ref)
            rot = U @ Vt
            J = ref[0, 0] * ref[1, 1]
            P = 2.0 * m['mu'] * (ref - rot) + m['lambda'] * (J - 1.0) * invRefMulDet.T
            pk[f] = inv @ P
            q1, q2 = Q[f, :, 0], Q[f, :, 1]
            f2 = -(P[0, 0] * inv[0, 0] + P[0, 1] * inv[0, 1]) * q1
            f3 = -P[0, 1] * inv[1, 1] * q1 - P[1, 1] * inv[1, 1] * q2
            a, b, c = m['faces'][f]
            vf[a] += -(f2 + f3); vf[b] += f2; vf[c] += f3
        return vf, pk

    # ------------------------------------------------------------------ HS:252-458
    def compute_grid_forces(self, Dt):
        rg = self.rg; rg.forces = np.zeros((rg.Ng, 3))
        if self.ps is not None:
            p = self.ps; n = p['x'].shape[0]
            lambda0 = p['E'] * p['nu'] / (1.0 + p['nu']) / (1.0 - 2.0 * p['nu']); mu0 = p['E'] / 2.0 / (1.0 + p['nu'])
            c1 = Dt * (self.dom1 @ rg.velocities); c2 = Dt * (self.dom2 @ rg.velocities); c3 = Dt * (self.dom3 @ rg.velocities)
            Mod = np.stack([c1, c2, c3], axis=2)                     # Mod[p][:, c] = row p of c_c
            FE = p['FE']; Fh = FE + Mod @ FE                          # HS:306
            U, s, Vt = np.linalg.svd(Fh)
            if self.material == SNOW:
                Jp = np.linalg.det(p['FP'])
                lam = lambda0 * np.exp(10.0 * (1 - Jp)); mu = mu0 * np.exp(10.0 * (1 - Jp))
                Rm = U @ Vt; J = np.linalg.det(Fh)
                P = 2.0 * mu[:, None, None] * (Fh - Rm) + (lam * (J - 1.0) * J)[:, None, None] * np.linalg.inv(np.swapaxes(Fh, 1, 2))
                stress = p['vol'][:, None, None] * (P @ np.swapaxes(FE, 1, 2))
            else:
                ls = np.log(s)
                dg = 2 * mu0 * (1.0 / s) * ls + lambda0 * ls.sum(axis=1)[:, None] * (1.0 / s)
                stress = p['vol'][:, None, None] * (((U * dg[:, None, :]) @ Vt) @ np.swapaxes(FE, 1, 2))
            for r in range(3):                                       # HS:356-366
                rg.forces[:, r] -= self.dom1.T @ stress[:, r, 0]
                rg.forces[:, r] -= self.dom2.T @ stress[:, r, 1]
                rg.forces[:, r] -= self.dom3.T @ stress[:, r, 2]
        if self.mesh is not None:
            m = self.mesh; nf = m['faces'].shape[0]
            vf, pk = self.cloth_in_plane()
            rg.forces = rg.forces + self.vom.T @ vf                  # HS:378
            Dm = np.stack([m['eD1'], m['eD2'], m['eD3']], axis=2); dm = np.stack([m['ed1'], m['ed2'], m['ed3']], axis=2)
            Q, R = gram_schmidt(dm)
            stress = np.zeros((nf, 3, 3))
            for f in range(nf):
                Rf = R[f]
                dR = np.array([[pk[f, 0, 0], pk[f, 0, 1], m['gamma'] * Rf[0, 2]],
                               [0.0, pk[f, 1, 1], m['gamma'] * Rf[1, 2]],
                               [0.0, 0.0, 0.0 if Rf[2, 2] > 1.0 else -m['k'] * (1.0 - Rf[2, 2]) * (1.0 - Rf[2, 2])]])
                K = dR @ Rf.T
                Sy = np.triu(K, 1) + np.triu(K).T                    # HS:425-426
                dF3 = Q[f] @ Sy @ np.linalg.inv(Rf).T @ Dm[f].T[:, 2]
                stress[f] = m['evol'][f] * np.outer(dF3, dm[f][:, 2])
            for r in range(3):                                       # HS:444-454
                rg.forces[:, r] -= self.deom1.T @ stress[:, r, 0]
                rg.forces[:, r] -= self.deom2.T @ stress[:, r, 1]
                rg.forces[:, r] -= self.deom3.T @ stress[:, r, 2]
        rg.forces[:, 2] -= rg.masses * 9.8                           # HS:457

    # ------------------------------------------------------------------ HS:725-737
    def update_grid_velocities(self, Dt):
        rg = self.rg; nz = rg.masses > 0.0
        rg.velocities[nz] += Dt * rg.forces[nz] / rg.masses[nz, None]

    # ------------------------------------------------------------------ HS:460-551
    def grid_collision(self):
        rg = self.rg; vbf = rg.velocities.copy(); friction = 0.2
        if self.phi is not None:
            inside = self.phi(rg.positions) <= 0.0
            ids = np.nonzero(inside)[0]
            if ids.size:
                nrm = self.dphi(rg.positions[ids]); vel = rg.velocities[ids]
                vn = (vel * nrm).sum(axis=1)
                app = vn < 0.0
                vt = vel - vn[:, None] * nrm
                vbf[ids[app]] = vt[app]
                stick = np.linalg.norm(vt, axis=1) < -friction * vn
                new = np.where(stick[:, None], 0.0, vt)              # HS:500-501 no-op friction reproduced
                rg.velocities[ids[app]] = new[app]
        if self.mesh is not None and self.mesh.get('fixed') is not None:
            coo = self.vom.tocoo(); nx, ny, nz_ = rg.res
            for vid, gid in zip(coo.row, coo.col):
                if self.mesh['fixed'][vid] == 0.0:
                    continue
                rk = gid // (nx * ny); rj = (gid % (nx * ny)) // nx; ri = gid - rk * nx * ny - rj * nx
                for i in range(ri - 1, ri + 2):
                    for j in range(rj - 1, rj + 2):
                        for k in range(rk - 1, rk + 2):
                            index = k * nx * ny + j * nx + i
                            if 0 <= index < rg.Ng:
                                rg.velocities[index] = 0.0; vbf[index] = 0.0
        return vbf

    # ------------------------------------------------------------------ HS:760-825
    def update_affine(self, om, pos, damp):
        rg = self.rg
        vt = om @ rg.velocities
        B = []
        for a in range(3):
            B.append(om @ (sp.diags(rg.velocities[:, a]) @ rg.positions) - vt[:, a][:, None] * pos)     # HS:797-806
        C = np.stack(B, axis=1)                                      # C[p][a][:] = row a
        sym = 0.5 * (C + np.swapaxes(C, 1, 2))
        C = (C - sym) + (1 - damp) * sym
        return C[:, 0, :].copy(), C[:, 1, :].copy(), C[:, 2, :].copy()

    # ------------------------------------------------------------------ HS:553-609, 612-723
    def update_deformation_and_plasticity(self, Dt, vbf):
        if self.ps is not None:
            p = self.ps
            Mod = np.stack([Dt * (self.dom1 @ vbf), Dt * (self.dom2 @ vbf), Dt * (self.dom3 @ vbf)], axis=2)
            cand = p['FE'] + Mod @ p['FE']
        if self.mesh is not None:
            m = self.mesh
            G = np.stack([self.deom1 @ vbf, self.deom2 @ vbf, self.deom3 @ vbf], axis=2)
            fa = m['faces']
            m['ed1'] = m['vx'][fa[:, 1]] - m['vx'][fa[:, 0]]
            m['ed2'] = m['vx'][fa[:, 2]] - m['vx'][fa[:, 0]]
            m['ed3'] = Dt * np.einsum('nij,nj->ni', G, m['ed3']) + m['ed3']
        if self.ps is not None:
            Ftot = cand @ p['FP']
            U, s, Vt = np.linalg.svd(cand)
            if self.material == SNOW:
                s = np.clip(s, 1.0 - p['thetaC'], 1.0 + p['thetaS'])
            else:
                lam = p['E'] * p['nu'] / (1.0 + p['nu']) / (1.0 - 2.0 * p['nu']); mu = p['E'] / 2.0 / (1.0 + p['nu'])
                q = p['q']
                phi = (35.0 + (9.0 * q - 10.0) * np.exp(-0.2 * q)) * np.pi / 180.0
                alpha = np.sqrt(2.0 / 3.0) * 2.0 * np.sin(phi) / (3.0 - np.sin(phi))
                ls = np.log(s); tr = ls.sum(axis=1)
                dv = ls - (tr / 3.0)[:, None]; dvn = np.linalg.norm(dv, axis=1)
                dg = dvn + (3.0 * lam + 2.0 * mu) / 2.0 / mu * tr * alpha
                case1 = dg <= 0.0
                case2 = (~case1) & ((dvn == 0.0) | (tr > 0.0))
                case3 = (~case1) & (~case2)
                s_new = s.copy(); q_new = q.copy()
                s_new[case2] = 1.0; q_new[case2] += np.linalg.norm(ls[case2], axis=1)
                with np.errstate(invalid='ignore', divide='ignore'):
                    Hp = ls - dg[:, None] * dv / dvn[:, None]
                s_new[case3] = np.exp(Hp[case3]); q_new[case3] += dg[case3]
                s = s_new; p['q'] = q_new
            V = np.swapaxes(Vt, 1, 2)
            p['FE'] = (U * s[:, None, :]) @ Vt
            p['FP'] = (V * (1.0 / s)[:, None, :]) @ np.swapaxes(U, 1, 2) @ Ftot
        if self.mesh is not None:
            m = self.mesh
            dm = np.stack([m['ed1'], m['ed2'], m['ed3']], axis=2)
            Q, R = gram_schmidt(dm)
            for f in range(dm.shape[0]):
                Rf = R[f]
                if Rf[2, 2] > 1.0:
                    Rf[2, 2] = 1.0; Rf[0, 2] = Rf[1, 2] = 0.0
                else:
                    fn = m['k'] * (Rf[2, 2] - 1.0) ** 2
                    fs = m['gamma'] * np.sqrt(Rf[0, 2] ** 2 + Rf[1, 2] ** 2)
                    if fs > m['cf'] * fn:
                        Rf[0, 2] *= m['cf'] * fn / fs; Rf[1, 2] *= m['cf'] * fn / fs
                m['ed3'][f] = Q[f] @ Rf[:, 2]

    def update_element_positions(self):
        m = self.mesh; fa = m['faces']
        m['ex'] = (m['vx'][fa[:, 0]] + m['vx'][fa[:, 1]] + m['vx'][fa[:, 2]]) / 3.0      # LM:371-380

    # ------------------------------------------------------------------ HS:827-1032
    def init(self):
        if self.mesh is not None:
            self.update_element_positions()
        self.rebuild_weights(); self.particle_to_grid(True)
        self.dt = self.cfl / max(3e2, self.rg.cfl_condition())

    def substep(self):
        rg = self.rg
        self.compute_grid_forces(self.dt)
        self.update_grid_velocities(self.dt)
        self.dt = self.cfl / max(3e2, rg.cfl_condition())
        if self.inner_t + self.dt >= 1.0 / 60.0:
            self.dt = 1.0 / 60.0 - self.inner_t; self.t += 1.0 / 60.0; self.inner_t = 0.0; self.frame_no += 1
        else:
            self.inner_t += self.dt
        vbf = self.grid_collision()
        self.vbf = vbf
        Dt = self.dt
        if self.ps is not None:
            self.ps['v'] = self.om @ rg.velocities                                        # HS:743
        if self.mesh is not None:
            m = self.mesh; fa = m['faces']
            m['vv'] = self.vom @ rg.velocities
            m['ev'] = (m['vv'][fa[:, 0]] + m['vv'][fa[:, 1]] + m['vv'][fa[:, 2]]) / 3.0
        if self.ps is not None:
            self.ps['B1'], self.ps['B2'], self.ps['B3'] = self.update_affine(self.om, self.ps['x'], 0.0)
        if self.mesh is not None:
            m = self.mesh
            m['vB1'], m['vB2'], m['vB3'] = self.update_affine(self.vom, m['vx'], 1.0)
            m['eB1'], m['eB2'], m['eB3'] = self.update_affine(self.eom, m['ex'], 1.0)
        if self.ps is not None:
            self.ps['x'] = self.om @ (rg.positions + Dt * vbf)                            # HS:944
        if self.mesh is not None:
            self.mesh['vx'] = self.vom @ (rg.positions + Dt * vbf)                        # HS:948
            self.update_element_positions()
        self.update_deformation_and_plasticity(Dt, vbf)
        self.rebuild_weights()
        self.particle_to_grid(False)
        return Dt


# ---------------------------------------------------------------------------------------------- scene glue
def levelset_callables(kind, P):
    """numpy versions of LS:8-42 (+ the sphere/box primitives of oracle/mpm_oracle.cpp)."""
    P = np.asarray(P, float)
    if kind == 1:
        return (lambda x: x[:, 2] - P[0]), (lambda x: np.tile([0.0, 0.0, 1.0], (x.shape[0], 1)))
    if kind == 2:
        def phi(x): return np.minimum(np.minimum(x[:, 2] - P[2], P[0] - x[:, 0]), P[1] - x[:, 1])
        def dphi(x):
            dz = np.abs(x[:, 2] - P[2]); dx = np.abs(P[0] - x[:, 0]); dy = np.abs(P[1] - x[:, 1])
            n = np.tile([-1.0, 0.0, 0.0], (x.shape[0], 1))
            n[dy <= dx] = [0.0, -1.0, 0.0]
            n[(dz <= dx) & (dz <= dy)] = [0.0, 0.0, 1.0]
            return n
        return phi, dphi
    if kind == 3:
        def phi(x): return np.minimum(np.linalg.norm(x - P[:3], axis=1) - P[3], x[:, 2] - P[4])
        def dphi(x):
            d = x - P[:3]; r = np.linalg.norm(d, axis=1)
            n = np.tile([0.0, 0.0, 1.0], (x.shape[0], 1))
            sph = (r - P[3] <= x[:, 2] - P[4]) & (r > 0)
            n[sph] = d[sph] / r[sph, None]
            return n
        return phi, dphi
    if kind == 4:
        N = np.array([[0, 0, 1], [0, 0, -1], [1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0]], float)
        def dist(x): return np.stack([x[:, 2] - P[2], P[5] - x[:, 2], x[:, 0] - P[0], P[3] - x[:, 0], x[:, 1] - P[1], P[4] - x[:, 1]], axis=1)
        return (lambda x: dist(x).min(axis=1)), (lambda x: N[np.argmin(dist(x), axis=1)])
    return None, None


def from_scene(scene):
    """Build a LiteralSolver from an anisotropicelastoplasticity_b200.scenes.Scene."""
    g = Grid(scene.grid.mn, scene.grid.mx, scene.grid.res)
    s = LiteralSolver(g, scene.material, scene.cfl)
    if scene.particles is not None:
        p = scene.particles
        s.ps = dict(x=p.x.copy(), v=p.v.copy(), B1=p.B[:, 0, :].copy(), B2=p.B[:, 1, :].copy(), B3=p.B[:, 2, :].copy(),
                    FE=p.FE.copy(), FP=p.FP.copy(), m=p.m.copy(), vol=p.vol.copy(), q=p.q.copy(), E=p.E, nu=p.nu,
                    thetaC=p.thetaC, thetaS=p.thetaS)
    if scene.mesh is not None:
        m = scene.mesh
        s.mesh = dict(vx=m.vx.copy(), vv=m.vv.copy(), vm=m.vm.copy(), vB1=m.vB[:, 0, :].copy(), vB2=m.vB[:, 1, :].copy(),
                      vB3=m.vB[:, 2, :].copy(), faces=m.faces.copy(), ev=m.ev.copy(), em=m.em.copy(), evol=m.evol.copy(),
                      eB1=m.eB[:, 0, :].copy(), eB2=m.eB[:, 1, :].copy(), eB3=m.eB[:, 2, :].copy(),
                      ed1=m.ed[0].copy(), ed2=m.ed[1].copy(), ed3=m.ed[2].copy(), eD1=m.eD[0].copy(), eD2=m.eD[1].copy(),
                      eD3=m.eD[2].copy(), fixed=None if m.fixed is None else m.fixed.copy(), mu=m.mu, k=m.stiff,
                      gamma=m.shear, cf=m.fric)
        s.mesh['lambda'] = m.lam
    s.phi, s.dphi = levelset_callables(scene.levelset.kind, scene.levelset.params)
    return s
