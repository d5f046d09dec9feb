"""Generate tests/golden/*.npz from the numpy/scipy literal transcription (oracle/literal_numpy.py).

Run:  python -m oracle.make_golden      (from the repo root; a few seconds)
These vectors come from the independent sparse-matrix transcription of the reference's algebra -- the SECOND pin of the C++
oracle; the first is the reference's own code (oracle/_ref, oracle/make_ref_golden.py -> tests/golden/ref_*.npz).
Each file holds the full initial state and the state after `nsteps` iterations of the HybridSolver.cpp:867-1032 loop.
"""
from __future__ import annotations

import os

import numpy as np

from anisotropicelastoplasticity_b200 import scenes as sc
from oracle import literal_numpy as ln

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def golden_scenes():
    rng = np.random.default_rng(11)
    out = {}
    out["sand_block"] = sc.small_block(material=sc.SAND, res=16, cells=3, seed=7)
    out["snow_block"] = sc.small_block(material=sc.SNOW, res=16, cells=3, seed=8)
    # boundary case: block touching the domain corner -> truncated stencils (HybridSolver.cpp:44-46)
    out["sand_corner"] = sc.small_block(material=sc.SAND, res=12, cells=2, seed=9, lo=(0.0, 0.0, 0.0), levelset=False)
    # cloth + sand coupling with pinned vertices, shear and friction active
    s = sc.small_block(material=sc.SAND, res=16, cells=3, seed=10, lo=(0.35, 0.35, 0.5))
    n = 7; edge = 0.4 / (n - 1)
    mesh = sc.make_cloth(n, n, (0.3, 0.3, 0.45), (edge, 0, 0), (0, edge, 0), shear=50.0, friction_angle_deg=20.0, fixed_ids=(0, n - 1))
    mesh.vx = mesh.vx + 0.01 * rng.standard_normal(mesh.vx.shape)
    mesh.vv = mesh.vv + 0.3 * rng.standard_normal(mesh.vv.shape)
    mesh.ed[0] = mesh.vx[mesh.faces[:, 1]] - mesh.vx[mesh.faces[:, 0]]
    mesh.ed[1] = mesh.vx[mesh.faces[:, 2]] - mesh.vx[mesh.faces[:, 0]]
    mesh.ed[2] = mesh.ed[2] * (1 + 0.05 * rng.standard_normal((mesh.nf, 1))) + 0.05 * rng.standard_normal((mesh.nf, 3))
    s.mesh = mesh
    s.levelset = sc.LevelSetSpec(sc.LS_GROUND, np.array([0.42, 0, 0, 0, 0, 0, 0, 0.0]))
    s.name = "cloth_sand"
    out["cloth_sand"] = s
    return out


def scene_to_dict(s: sc.Scene):
    d = dict(grid_mn=s.grid.mn, grid_mx=s.grid.mx, grid_res=s.grid.res, material=s.material, cfl=s.cfl,
             ls_kind=s.levelset.kind, ls_params=s.levelset.params)
    if s.particles is not None:
        p = s.particles
        d.update(p_x=p.x, p_v=p.v, p_B=p.B, p_FE=p.FE, p_FP=p.FP, p_m=p.m, p_vol=p.vol, p_q=p.q,
                 p_const=np.array([p.E, p.nu, p.thetaC, p.thetaS]))
    if s.mesh is not None:
        m = s.mesh
        d.update(m_vx=m.vx, m_vv=m.vv, m_vm=m.vm, m_vvol=m.vvol, m_vB=m.vB, m_faces=m.faces, m_ev=m.ev, m_em=m.em, m_evol=m.evol,
                 m_eB=m.eB, m_ed=m.ed, m_eD=m.eD, m_fixed=(np.zeros(0) if m.fixed is None else m.fixed),
                 m_const=np.array([m.mu, m.lam, m.shear, m.stiff, m.fric]))
    return d


def scene_from_dict(d, name="golden"):
    g = sc.GridSpec(d["grid_mn"], d["grid_mx"], d["grid_res"])
    ps = None; mesh = None
    if "p_x" in d:
        c = d["p_const"]
        ps = sc.Particles(x=d["p_x"].copy(), v=d["p_v"].copy(), B=d["p_B"].copy(), FE=d["p_FE"].copy(), FP=d["p_FP"].copy(), m=d["p_m"].copy(),
                          vol=d["p_vol"].copy(), q=d["p_q"].copy(), E=float(c[0]), nu=float(c[1]), thetaC=float(c[2]), thetaS=float(c[3]))
    if "m_vx" in d:
        c = d["m_const"]
        mesh = sc.Mesh(vx=d["m_vx"].copy(), vv=d["m_vv"].copy(), vm=d["m_vm"].copy(), vvol=d["m_vvol"].copy(), vB=d["m_vB"].copy(),
                       faces=d["m_faces"].copy(), ev=d["m_ev"].copy(), em=d["m_em"].copy(), evol=d["m_evol"].copy(), eB=d["m_eB"].copy(),
                       ed=d["m_ed"].copy(), eD=d["m_eD"].copy(), fixed=(d["m_fixed"].copy() if d["m_fixed"].size else None),
                       mu=float(c[0]), lam=float(c[1]), shear=float(c[2]), stiff=float(c[3]), fric=float(c[4]))
    ls = sc.LevelSetSpec(int(d["ls_kind"]), np.asarray(d["ls_params"], float))
    return sc.Scene(name, g, int(d["material"]), ps, mesh, ls, float(d["cfl"]))


def run_literal(scene, nsteps):
    l = ln.from_scene(scene); l.init()
    out = dict(dt0=l.dt)
    if l.ps is not None:
        out["vol_init"] = l.ps["vol"].copy()
    out["g0_m"] = l.rg.masses.copy(); out["g0_v"] = l.rg.velocities.copy()
    dts = []
    for _ in range(nsteps):
        dts.append(l.substep())
    out["dts"] = np.array(dts)
    if l.ps is not None:
        p = l.ps
        out.update(o_x=p["x"], o_v=p["v"], o_B=np.stack([p["B1"], p["B2"], p["B3"]], axis=1), o_FE=p["FE"], o_FP=p["FP"], o_q=p["q"])
    if l.mesh is not None:
        m = l.mesh
        out.update(o_vx=m["vx"], o_vv=m["vv"], o_vB=np.stack([m["vB1"], m["vB2"], m["vB3"]], axis=1), o_ex=m["ex"], o_ev=m["ev"],
                   o_eB=np.stack([m["eB1"], m["eB2"], m["eB3"]], axis=1), o_ed=np.stack([m["ed1"], m["ed2"], m["ed3"]], axis=0))
    out.update(o_gm=l.rg.masses, o_gv=l.rg.velocities, o_gf=l.rg.forces, o_vbf=l.vbf)
    return out


def main(nsteps=3):
    os.makedirs(OUT, exist_ok=True)
    for name, scene in golden_scenes().items():
        d = scene_to_dict(scene)
        d.update(run_literal(scene, nsteps)); d["nsteps"] = nsteps
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **d)
        print(name, "->", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
