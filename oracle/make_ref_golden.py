"""Generate tests/golden/ref_*.npz by running the REFERENCE'S OWN CODE (oracle/_ref/libaep_ref.so: the unmodified sources
under /root/reference compiled against the MiniEigen stand-in, oracle/Makefile) on seeded small scenes.

Run:  python -m oracle.make_ref_golden      (from the repo root, in the dev container where /root/reference exists; ~10 s)

These fixtures are what pins the oracle (tests/test_reference_pin.py, CPU) and, directly, the CUDA engine
(tests/test_gpu_parity.py::test_reference_golden_substeps) to the reference where /root/reference does not exist (the GPU
box).  Same keys as oracle/make_golden.py's files: full initial state + state after `nsteps` passes of HybridSolver.cpp:867-1032.
`ref_solve_frame.npz` is different: its output is what HybridSolver::solve ITSELF leaves behind after one 1/60 s frame,
including the particle_0.obj it writes.
"""
from __future__ import annotations

import os
import tempfile

import numpy as np

from anisotropicelastoplasticity_b200 import scenes as sc
from oracle.make_golden import OUT, golden_scenes, scene_to_dict
from oracle.ref_py import Reference


def ref_scenes():
    out = dict(golden_scenes())                                     # sand_block, snow_block, sand_corner, cloth_sand
    # sand resting against both walls and the ground of the wall-corner collider: sticking and sliding nodes (HS:486-507)
    s = sc.small_block(material=sc.SAND, res=16, cells=3, seed=21, lo=(0.4375, 0.4375, 0.25))
    s.levelset = sc.LevelSetSpec(sc.LS_WALL2GROUND, np.array([0.62, 0.62, 0.27, 0, 0, 0, 0, 0.0])); s.name = "sand_walls"
    out["sand_walls"] = s
    # snow falling onto a sphere + ground: a collider the reference lacks, fed to its collision code as nodal samples
    s = sc.small_block(material=sc.SNOW, res=16, cells=3, seed=22, lo=(0.375, 0.375, 0.3125))
    s.particles.v[:, 2] -= 2.0
    s.levelset = sc.LevelSetSpec(sc.LS_SPHERE_GROUND, np.array([0.5, 0.5, 0.22, 0.12, 0.1, 0, 0, 0.0])); s.name = "snow_sphere"
    out["snow_sphere"] = s
    # cloth alone with two pinned corners over the ground: the configuration main.cpp:82-84 actually runs
    n = 9; edge = 0.5 / (n - 1); rng = np.random.default_rng(23)
    mesh = sc.make_cloth(n, n, (0.25, 0.25, 0.5), (edge, 0, 0), (0, edge, 0), shear=30.0, friction_angle_deg=15.0, fixed_ids=(0, n - 1))
    mesh.vx = mesh.vx + 0.004 * rng.standard_normal(mesh.vx.shape); mesh.vv = mesh.vv + 0.2 * rng.standard_normal(mesh.vv.shape)
    mesh.ed[0] = mesh.vx[mesh.faces[:, 1]] - mesh.vx[mesh.faces[:, 0]]; mesh.ed[1] = mesh.vx[mesh.faces[:, 2]] - mesh.vx[mesh.faces[:, 0]]
    g = sc.GridSpec(np.zeros(3), np.ones(3), np.array([16, 16, 16]))
    out["cloth_only"] = sc.Scene("cloth_only", g, sc.SAND, None, mesh, sc.LevelSetSpec(sc.LS_GROUND, np.array([0.3, 0, 0, 0, 0, 0, 0, 0.0])))
    return out


def state(r: Reference, scene, prefix="o_"):
    out = {}
    if scene.particles is not None:
        p = r.particles()
        out.update({prefix + k: p[k] for k in ("x", "v", "B", "FE", "FP", "q")})
    if scene.mesh is not None:
        m = r.mesh()
        out.update({prefix + k: m[k] for k in ("vx", "vv", "vB", "ex", "ev", "eB", "ed")})
    g = r.grid()
    out.update({prefix + "gm": g["m"], prefix + "gv": g["v"], prefix + "gf": g["f"], prefix + "vbf": g["vt"]})
    return out


def run_reference(scene, nsteps):
    r = Reference(scene); r.init()
    out = dict(dt0=r.dt)
    if scene.particles is not None:
        out["vol_init"] = r.particles()["vol"].copy()
    g = r.grid(); out["g0_m"] = g["m"].copy(); out["g0_v"] = g["v"].copy()
    out["dts"] = np.array([r.substep() for _ in range(nsteps)])
    out.update(state(r, scene))
    return out


def run_solve_frame():
    """HybridSolver::solve(0.3, 0.0, 0.95): `while (t <= maxt)` ends when the first frame completes (HS:867,883)."""
    scene = sc.small_block(material=sc.SAND, res=16, cells=3, seed=31, lo=(0.3125, 0.3125, 0.3125)); scene.name = "solve_frame"
    r = Reference(scene)
    with tempfile.TemporaryDirectory() as tmp:
        r.solve(0.0, tmp)
        obj = open(os.path.join(tmp, "particle", "particle_0.obj")).read()
        assert not os.path.exists(os.path.join(tmp, "particle", "particle_1.obj"))
    # solve() cannot report its time steps, so the same scene goes through ref_driver.cpp's loop (the reference's stage
    # methods + the restated glue of HS:878-892) until the frame counter ticks: the two must agree BIT FOR BIT, which
    # shows the glue is the reference's, and yields the dt sequence for tests that replay it.
    r2 = Reference(scene); r2.init(); dt0 = r2.dt; dts = []
    while r2.frame == 0:
        dts.append(r2.substep())
    a, b = state(r, scene), state(r2, scene)
    # (gridVelocityBeforeFriction is a local of solve(), HS:897: only the driver's run can report it)
    assert all(np.array_equal(a[k], b[k]) for k in a if k != "o_vbf"), "ref_driver's loop differs from HybridSolver::solve"
    a["o_vbf"] = b["o_vbf"]
    d = scene_to_dict(scene); d.update(a); d["dt0"] = dt0; d["dts"] = np.array(dts)
    d["obj_head"] = np.array(obj.splitlines()[:8])
    d["obj_xyz"] = np.array([[float(t) for t in ln.split()[1:]] for ln in obj.splitlines() if ln.startswith("v ")])
    return d


def main(nsteps=6):
    os.makedirs(OUT, exist_ok=True)
    for name, scene in ref_scenes().items():
        d = scene_to_dict(scene); d.update(run_reference(scene, nsteps)); d["nsteps"] = nsteps
        path = os.path.join(OUT, "ref_" + name + ".npz"); np.savez_compressed(path, **d)
        print(name, "->", path, os.path.getsize(path) // 1024, "KiB")
    path = os.path.join(OUT, "ref_solve_frame.npz"); np.savez_compressed(path, **run_solve_frame())
    print("solve_frame ->", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
