"""Seeded synthetic scenes for the MPM substep hot path (BASELINE.json `configs`).

The reference seeds its factories from ``time(0)`` / ``rand()`` (ParticleSystem.cpp:129,189,258,338,344) and its
one mesh asset is missing (main.cpp:78), so nothing it builds is reproducible.  These generators replace them with
``numpy.random.default_rng(seed)`` jittered lattices at 8 particles per cell, keeping the reference's material
constants (ParticleSystem.cpp:173-177, 233-234, 321-322; main.cpp:77-78).

Pure numpy; shared by tests, bench.py and the engine wrappers.  Array conventions (host side, float64):
    x, v            (N, 3)
    B, FE, FP       (N, 3, 3)   B[p][a] is row a == affineMomenta_{a+1}.row(p)
    m, vol, q       (N,)
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np

SNOW, SAND = 0, 1                       # HybridSolver.h:21-25
LS_NONE, LS_GROUND, LS_WALL2GROUND, LS_SPHERE_GROUND, LS_BOX, LS_SAMPLED = 0, 1, 2, 3, 4, 5


@dataclass
class GridSpec:
    mn: np.ndarray
    mx: np.ndarray
    res: np.ndarray

    @property
    def h(self):
        return (np.asarray(self.mx, float) - np.asarray(self.mn, float)) / np.asarray(self.res)     # RegularGrid.cpp:137-139

    @property
    def n_nodes(self):
        return int(np.prod(self.res))


@dataclass
class Particles:
    x: np.ndarray
    v: np.ndarray
    B: np.ndarray
    FE: np.ndarray
    FP: np.ndarray
    m: np.ndarray
    vol: np.ndarray
    q: np.ndarray
    E: float
    nu: float
    thetaC: float = 2.5e-2
    thetaS: float = 7.5e-3

    @property
    def n(self):
        return self.x.shape[0]


@dataclass
class Mesh:
    """LagrangianMesh public state (LagrangianMesh.h:38-74)."""
    vx: np.ndarray          # (Nv,3) vertexPositions
    vv: np.ndarray          # (Nv,3)
    vm: np.ndarray          # (Nv,)
    vvol: np.ndarray
    vB: np.ndarray          # (Nv,3,3)
    faces: np.ndarray       # (Nf,3) int32
    ev: np.ndarray          # (Nf,3) elementVelocities
    em: np.ndarray
    evol: np.ndarray
    eB: np.ndarray          # (Nf,3,3)
    ed: np.ndarray          # (3,Nf,3) elementDirections_{1,2,3}
    eD: np.ndarray          # (3,Nf,3) rest directions
    fixed: Optional[np.ndarray]
    mu: float
    lam: float
    shear: float            # shearStiffness (gamma)
    stiff: float            # stiffness (k)
    fric: float             # frictionCoeff = tan(angle)  (LagrangianMesh.cpp:351)

    @property
    def nv(self):
        return self.vx.shape[0]

    @property
    def nf(self):
        return self.faces.shape[0]


@dataclass
class LevelSetSpec:
    kind: int = LS_NONE
    params: np.ndarray = field(default_factory=lambda: np.zeros(8))


@dataclass
class Scene:
    name: str
    grid: GridSpec
    material: int
    particles: Optional[Particles]
    mesh: Optional[Mesh] = None
    levelset: LevelSetSpec = field(default_factory=LevelSetSpec)
    cfl: float = 0.3                    # main.cpp:27


# ---------------------------------------------------------------------------------------------- sampling
def jittered_lattice(lo, hi, h, rng, ppc_axis=2, keep=None, dtype=np.float64):
    """ppc_axis^3 particles per cell, one per sub-cell, uniformly jittered inside the sub-cell.

    lo/hi are snapped outward to cell boundaries of a lattice with spacing h anchored at the origin."""
    lo = np.asarray(lo, float); hi = np.asarray(hi, float); h = np.asarray(h, float)
    c0 = np.floor(lo / h + 1e-9).astype(np.int64); c1 = np.ceil(hi / h - 1e-9).astype(np.int64)
    n = (c1 - c0) * ppc_axis
    sub = h / ppc_axis
    out = []
    # chunk over z to bound temporary memory
    zchunk = max(1, int(4_000_000 // max(1, n[0] * n[1])))
    for z0 in range(0, int(n[2]), zchunk):
        z1 = min(int(n[2]), z0 + zchunk)
        k, j, i = np.meshgrid(np.arange(z0, z1), np.arange(n[1]), np.arange(n[0]), indexing='ij')
        ijk = np.stack([i.ravel(), j.ravel(), k.ravel()], axis=1).astype(np.float64)
        jit = rng.random(ijk.shape)
        p = c0 * h + (ijk + 0.05 + 0.9 * jit) * sub
        if keep is not None:
            p = p[keep(p)]
        out.append(p.astype(dtype, copy=False))
    return np.concatenate(out, axis=0)


def _lame(E, nu):
    return E * nu / (1.0 + nu) / (1.0 - 2.0 * nu), E / 2.0 / (1.0 + nu)


def make_particles(x, density, h, E, nu, v0=(0.0, 0.0, 0.0), ppc=8, thetaC=2.5e-2, thetaS=7.5e-3):
    n = x.shape[0]
    eye = np.broadcast_to(np.eye(3), (n, 3, 3)).copy()
    m = np.full(n, density * float(np.prod(h)) / ppc)
    return Particles(x=x, v=np.broadcast_to(np.asarray(v0, float), (n, 3)).copy(), B=np.zeros((n, 3, 3)),
                     FE=eye, FP=eye.copy(), m=m, vol=np.ones(n), q=np.zeros(n), E=E, nu=nu, thetaC=thetaC, thetaS=thetaS)


def perturb_state(ps: Particles, rng, strain=2e-2, vel=0.5, affine=2.0):
    """Non-trivial F_E, F_P, v, B, q so that a single substep exercises every branch (parity tests)."""
    n = ps.n
    ps.FE = ps.FE + strain * rng.standard_normal((n, 3, 3))
    ps.FP = ps.FP + 0.5 * strain * rng.standard_normal((n, 3, 3))
    ps.v = ps.v + vel * rng.standard_normal((n, 3))
    ps.B = affine * rng.standard_normal((n, 3, 3))
    ps.q = np.abs(0.3 * rng.standard_normal(n))
    return ps


# ---------------------------------------------------------------------------------------------- cloth
def make_cloth(nu_, nv_, origin, du, dv, density=2e3, thickness=0.04, E=200.0, nu=0.3, shear=0.0, stiff=4e4,
               friction_angle_deg=0.0, fixed_ids=()):
    """Regular nu_ x nv_ vertex sheet, two triangles per quad, with the volume / mass / rest-direction rules of
    LagrangianMesh::ObjMesh (LagrangianMesh.cpp:306-351) and main.cpp:77-78 material parameters."""
    origin = np.asarray(origin, float); du = np.asarray(du, float); dv = np.asarray(dv, float)
    iu, iv = np.meshgrid(np.arange(nu_), np.arange(nv_), indexing='ij')
    V = origin + iu.reshape(-1, 1) * du + iv.reshape(-1, 1) * dv
    vid = lambda a, b: a * nv_ + b
    a, b = np.meshgrid(np.arange(nu_ - 1), np.arange(nv_ - 1), indexing='ij'); a = a.ravel(); b = b.ravel()
    f1 = np.stack([vid(a, b), vid(a + 1, b), vid(a + 1, b + 1)], axis=1)
    f2 = np.stack([vid(a, b), vid(a + 1, b + 1), vid(a, b + 1)], axis=1)
    F = np.concatenate([f1, f2], axis=0).astype(np.int32)
    v1, v2, v3 = V[F[:, 0]], V[F[:, 1]], V[F[:, 2]]
    la = np.linalg.norm(v2 - v1, axis=1); lb = np.linalg.norm(v3 - v2, axis=1); lc = np.linalg.norm(v1 - v3, axis=1)
    s = 0.5 * (la + lb + lc)
    area = np.sqrt(np.maximum(s * (s - la) * (s - lb) * (s - lc), 0.0))          # geometry.cpp:13-24 (Heron)
    evol = 0.25 * area * thickness                                                # LagrangianMesh.cpp:315
    vvol = np.zeros(V.shape[0])
    for c in range(3):
        np.add.at(vvol, F[:, c], evol)                                            # LagrangianMesh.cpp:318-320
    nrm = np.cross(v2 - v1, v3 - v1); nrm /= np.linalg.norm(nrm, axis=1)[:, None]
    eD = np.stack([v2 - v1, v3 - v1, nrm], axis=0)                                # LagrangianMesh.cpp:324-326
    lam, mu = _lame(E, nu)
    fixed = None
    if len(fixed_ids):
        fixed = np.zeros(V.shape[0]); fixed[list(fixed_ids)] = 1.0
    nv, nf = V.shape[0], F.shape[0]
    return Mesh(vx=V, vv=np.zeros((nv, 3)), vm=density * vvol, vvol=vvol, vB=np.zeros((nv, 3, 3)), faces=F,
                ev=np.zeros((nf, 3)), em=density * evol, evol=evol, eB=np.zeros((nf, 3, 3)), ed=eD.copy(), eD=eD,
                fixed=fixed, mu=mu, lam=lam, shear=shear, stiff=stiff,
                fric=float(np.tan(friction_angle_deg * np.pi / 180.0)))


# ---------------------------------------------------------------------------------------------- configs
SAND_E, SAND_NU, SAND_RHO = 3.537e5, 0.3, 1300.0          # ParticleSystem.cpp:233-234,296,321-322
SNOW_E, SNOW_NU, SNOW_RHO = 1.4e5, 0.2, 400.0             # ParticleSystem.cpp:173-177


def c1_sand_block(res=64, seed=1):
    """C1: sand block with a spherical hole dropped onto a wall-corner collider (SandBlock analogue,
    ParticleSystem.cpp:240-327), ~1e5 particles on a 64^3 grid at res=64."""
    rng = np.random.default_rng(seed)
    g = GridSpec(np.zeros(3), np.ones(3), np.array([res] * 3))
    lo = np.array([0.30, 0.30, 0.15]); hi = np.array([0.60, 0.60, 0.65])
    hole_c = np.array([lo[0], lo[1], 0.5 * (lo[2] + hi[2])])                      # ParticleSystem.cpp:251
    keep = lambda p: np.linalg.norm(p - hole_c, axis=1) >= 0.08
    x = jittered_lattice(lo, hi, g.h, rng, keep=keep)
    ps = make_particles(x, SAND_RHO, g.h, SAND_E, SAND_NU)
    ls = LevelSetSpec(LS_WALL2GROUND, np.array([0.9, 0.9, 0.1, 0, 0, 0, 0, 0.0]))
    return Scene("C1_sand_block", g, SAND, ps, None, ls)


def c2_snow_sphere(res=128, seed=2):
    """C2: snow block thrown at a level-set sphere resting on the ground, ~1e6 particles on 128^3 at res=128."""
    rng = np.random.default_rng(seed)
    g = GridSpec(np.zeros(3), np.ones(3), np.array([res] * 3))
    side = 50.0 / 128.0
    lo = np.array([0.5 - side / 2, 0.5 - side / 2, 0.40]); hi = lo + side
    x = jittered_lattice(lo, hi, g.h, rng)
    ps = make_particles(x, SNOW_RHO, g.h, SNOW_E, SNOW_NU, v0=(0.0, 0.0, -3.0))
    ls = LevelSetSpec(LS_SPHERE_GROUND, np.array([0.5, 0.5, 0.2, 0.12, 0.05 + 1e-4, 0, 0, 0.0]))
    return Scene("C2_snow_sphere", g, SNOW, ps, None, ls)


def c5_dam_break(res=512, seed=5, slab=None):
    """C5: sand dam-break column in a box, ~6.4e7 particles on 512^3 at res=512 (scales as res^3)."""
    rng = np.random.default_rng(seed)
    g = GridSpec(np.zeros(3), np.ones(3), np.array([res] * 3))
    h = g.h
    lo = np.array([2 * h[0], 2 * h[1], 2 * h[2]]); hi = np.array([0.25, 1.0 - 2 * h[1], 0.25])
    x = jittered_lattice(lo, hi, h, rng, dtype=np.float64)
    ps = make_particles(x, SAND_RHO, h, SAND_E, SAND_NU)
    e = 1e-4 * h[0]
    ls = LevelSetSpec(LS_BOX, np.array([2 * h[0] - e, 2 * h[1] - e, 2 * h[2] - e, 1 - 2 * h[0] + e, 1 - 2 * h[1] + e, 1 - 2 * h[2] + e, 0, 0.0]))
    return Scene("C5_dam_break", g, SAND, ps, None, ls)


def c3_cloth_drape(n=512, seed=3, grid_h=None):
    """C3: n x n cloth draping over a sphere + ground (main.cpp:51-91 analogue with a sphere collider)."""
    edge = 1.0 / (n - 1)
    h = edge if grid_h is None else grid_h                                        # main.cpp:53-54
    mn = np.array([-0.1, -0.1, 0.0]); mx = np.array([1.1, 1.1, 0.8])
    res = np.maximum(1, np.floor((mx - mn) / h + 0.5).astype(int))                # main.cpp:63-65
    g = GridSpec(mn, mx, res)
    mesh = make_cloth(n, n, (0.0, 0.0, 0.7), (edge, 0, 0), (0, edge, 0))
    ls = LevelSetSpec(LS_SPHERE_GROUND, np.array([0.5, 0.5, 0.4, 0.25, 0.05 + 1e-4, 0, 0, 0.0]))
    return Scene("C3_cloth_drape", g, SAND, None, mesh, ls)


def c4_coupling(res=256, cloth_n=256, seed=4):
    """C4: sand dropped onto a cloth pinned at two corners (video/coupling.mp4 analogue), ~4e6 particles at res=256."""
    rng = np.random.default_rng(seed)
    g = GridSpec(np.zeros(3), np.ones(3), np.array([res] * 3))
    lo = np.array([0.5 - 40.0 / 256, 0.5 - 40.0 / 256, 0.55]); hi = np.array([0.5 + 40.0 / 256, 0.5 + 40.0 / 256, 0.55 + 78.0 / 256])
    x = jittered_lattice(lo, hi, g.h, rng)
    ps = make_particles(x, SAND_RHO, g.h, SAND_E, SAND_NU)
    edge = 0.6 / (cloth_n - 1)
    mesh = make_cloth(cloth_n, cloth_n, (0.2, 0.2, 0.5), (edge, 0, 0), (0, edge, 0), fixed_ids=(0, cloth_n - 1))
    ls = LevelSetSpec(LS_GROUND, np.array([0.05 + 1e-4, 0, 0, 0, 0, 0, 0, 0.0]))
    return Scene("C4_coupling", g, SAND, ps, mesh, ls)


def small_block(material=SAND, res=16, seed=7, cells=5, perturb=True, levelset=True, lo=(0.3, 0.3, 0.25)):
    """Tiny parity scene: cells^3 cells of particles (8/cell), optional perturbed state and wall-corner collider."""
    rng = np.random.default_rng(seed)
    g = GridSpec(np.zeros(3), np.ones(3), np.array([res] * 3))
    lo = np.asarray(lo, float); hi = lo + cells * g.h
    x = jittered_lattice(lo, hi, g.h, rng)
    if material == SAND:
        ps = make_particles(x, SAND_RHO, g.h, SAND_E, SAND_NU)
    else:
        ps = make_particles(x, SNOW_RHO, g.h, SNOW_E, SNOW_NU, v0=(0.0, 0.0, -1.0))
    if perturb:
        perturb_state(ps, rng)
    ls = LevelSetSpec(LS_WALL2GROUND, np.array([0.66, 0.66, 0.27, 0, 0, 0, 0, 0.0])) if levelset else LevelSetSpec()
    return Scene("small_block", g, material, ps, None, ls)


CONFIGS = {"C1": c1_sand_block, "C2": c2_snow_sphere, "C3": c3_cloth_drape, "C4": c4_coupling, "C5": c5_dam_break}


# ---------------------------------------------------------------------------------------------- layout helpers
def colmajor(a):
    """(N,3) -> Eigen MatrixX3d memory (column-major, ld = N)."""
    return np.ascontiguousarray(np.asarray(a, np.float64).T)


def from_colmajor(buf, n):
    return np.ascontiguousarray(buf.reshape(3, n).T)


def mats_colmajor(a):
    """(N,3,3) [p][r][c] -> std::vector<Matrix3d> memory (each 3x3 column-major)."""
    return np.ascontiguousarray(np.asarray(a, np.float64).transpose(0, 2, 1))


def mats_from_colmajor(buf, n):
    return np.ascontiguousarray(buf.reshape(n, 3, 3).transpose(0, 2, 1))


def bulk_stats(x, v, m, FP):
    """Bulk statistics used for the 200-substep gate (BASELINE.json north_star)."""
    M = m.sum()
    com = (x * m[:, None]).sum(axis=0) / M
    ke = 0.5 * (m * (v * v).sum(axis=1)).sum()
    jp = np.linalg.det(FP).mean()
    return com, ke, jp
