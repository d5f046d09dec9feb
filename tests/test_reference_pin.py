"""Pin the oracle to the REFERENCE'S OWN CODE.  CPU only.

oracle/_ref/libaep_ref.so is the reference's unmodified C++ (compiled where it lies under /root/reference against the
MiniEigen stand-in of oracle/ref_shim; oracle/Makefile, oracle/ref_driver.cpp).  Two layers:

* fixture tests (always run): tests/golden/ref_*.npz were written by that library (oracle/make_ref_golden.py); the oracle
  restatement must reproduce them to fp64 rounding, including what HybridSolver::solve itself leaves after one frame.
* live tests (run wherever the library is built or buildable -- the dev container, and the GPU box, which receives the
  built file): the reference's scalar kernels, its stage methods from identical states on fresh seeds, the fixtures'
  provenance, and the stand-in's own SVD / sparse algebra against numpy / scipy.
"""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import load_golden, relerr
from oracle import ref_py
from oracle.oracle_py import Oracle
from oracle import oracle_py

live = pytest.mark.skipif(not ref_py.available(), reason="oracle/_ref/libaep_ref.so not built and /root/reference absent")

REF_GOLDEN = ["sand_block", "snow_block", "sand_corner", "cloth_sand", "sand_walls", "snow_sphere", "cloth_only"]
# The oracle and the reference are both fp64 but sum in different orders (direct stencil vs sparse column products), and
# the reference's "optimised" APIC algebra cancels x_i*sum(w m B) against sum(w m B x_p) (HS:178-203): momenta agree to
# ~1e-12, velocities at nodes of vanishing mass only to ~1e-7.
TOL, MTOL, VTOL = 1e-10, 1e-10, 1e-5


def mom(m, v):
    return np.asarray(m)[:, None] * np.asarray(v)


def compare_states(o, scene, d, tol=TOL):
    g = o.grid()
    assert relerr(g["m"], d["o_gm"]) < tol and relerr(mom(g["m"], g["v"]), mom(d["o_gm"], d["o_gv"])) < max(MTOL, tol)
    assert relerr(g["v"], d["o_gv"]) < VTOL and relerr(g["f"], d["o_gf"]) < tol and relerr(g["vt"], d["o_vbf"]) < VTOL
    if scene.particles is not None:
        p = o.particles()
        for k in ("x", "v", "B", "FE", "FP", "q"):
            assert relerr(p[k], d["o_" + k]) < tol, k
    if scene.mesh is not None:
        m = o.mesh()
        for k in ("vx", "vv", "vB", "ex", "ev", "eB", "ed"):
            assert relerr(m[k], d["o_" + k]) < tol, k


# ------------------------------------------------------------------------------------------------ fixtures (always)
@pytest.mark.parametrize("name", REF_GOLDEN)
def test_oracle_reproduces_reference_fixture(name):
    d, scene = load_golden("ref_" + name)
    o = Oracle(scene); o.init()
    assert o.dt == pytest.approx(float(d["dt0"]), rel=1e-12)                       # HS:860
    g0 = o.grid()
    assert relerr(g0["m"], d["g0_m"]) < 1e-13 and relerr(mom(g0["m"], g0["v"]), mom(d["g0_m"], d["g0_v"])) < MTOL
    if scene.particles is not None:
        assert relerr(o.particles()["vol"], d["vol_init"]) < 1e-13               # HS:242-249
    # Replay the reference's recorded time steps so that the states can be held to fp64 rounding: the dt rule itself takes
    # max |v_i| over ALL nodes with m > 0 (RegularGrid.cpp:188-200), including nodes of vanishing mass where the reference's
    # APIC algebra cancels catastrophically (HS:178-203), so the rule's output is only reproducible to ~1e-7 and is checked
    # separately, loosely.
    replay(o, scene, d["dt0"], d["dts"])
    compare_states(o, scene, d)
    # and free-running (the oracle's own dt rule and lag, orc_substep)
    f = Oracle(scene); f.init()
    assert np.allclose([f.substep() for _ in range(int(d["nsteps"]))], d["dts"], rtol=1e-5)
    compare_states(f, scene, d, tol=1e-6)


def replay(o, scene, dt0, dts, rule_rel=1e-5, frame_dt=1.0 / 60.0):
    """Drive the oracle's stages with the reference's recorded time steps (forces and the grid update use the LAGGED one,
    HS:873-877) and check the dt rule HS:878-892 at every step against what the reference chose."""
    dt_prev = float(dt0); inner = 0.0
    for dt in map(float, dts):
        o.stage_forces(dt_prev); o.stage_grid_update(dt_prev)
        rule = scene.cfl / max(3e2, o.cfl_condition())                              # HS:878
        if inner + rule >= frame_dt:                                                # HS:880-888: clipped to the frame
            assert dt == pytest.approx(frame_dt - inner, rel=1e-9, abs=1e-15); inner = 0.0
        else:
            assert rule == pytest.approx(dt, rel=rule_rel); inner += dt
        o.stage_collide(); o.stage_g2p(dt); o.rebuild_weights(); o.p2g(False)
        dt_prev = dt


def test_oracle_reproduces_reference_solve_frame():
    """HybridSolver::solve(0.3, 0, 0.95) ran one whole frame by itself (its own loop, dt lag, frame clipping, OBJ writer);
    oracle/make_ref_golden.py also checked that ref_driver.cpp's loop lands on the same state BIT FOR BIT and recorded its
    64 time steps.  Free-running, the oracle cannot follow that trajectory beyond ~1e-3: the dt rule amplifies rounding
    (it takes max|v| over nodes of vanishing mass, where v = p/m is noise) and dt swings by 100x between substeps.  With
    the recorded steps replayed it must land on the reference's state to rounding."""
    d, scene = load_golden("ref_solve_frame")
    assert len(d["dts"]) >= 17 and float(np.sum(d["dts"])) == pytest.approx(1.0 / 60.0, rel=1e-12)    # dt <= 0.3/300, one 1/60 s frame
    o = Oracle(scene); o.init()
    assert o.dt == pytest.approx(float(d["dt0"]), rel=1e-12)
    replay(o, scene, d["dt0"], d["dts"], rule_rel=1e-4)
    compare_states(o, scene, d, tol=1e-9)
    # particle_0.obj as the reference wrote it: "v x y z" at ostream's default 6 significant digits (HS:1000-1006)
    assert [str(s) for s in d["obj_head"]][0].startswith("v ")
    assert np.allclose(d["obj_xyz"], o.particles()["x"], rtol=2e-5, atol=0)
    # the oracle's own loop (orc_substep) implements the same frame logic: it also needs 1/60 s to tick, within the chaos band
    f = Oracle(scene); f.init(); n = 0
    while f.frame == 0:
        f.substep(); n += 1
        assert n < 1000
    assert f.time == pytest.approx(1.0 / 60.0, rel=1e-14) and relerr(f.particles()["x"], d["o_x"]) < 2e-2


# ------------------------------------------------------------------------------------------------ live
@live
def test_fixtures_come_from_the_reference_build():
    from oracle.make_ref_golden import ref_scenes, run_reference
    for name in ("sand_walls", "cloth_sand"):
        d, _ = load_golden("ref_" + name)
        out = run_reference(ref_scenes()[name], int(d["nsteps"]))
        for k in ("dts", "o_gm", "o_gf") + (("o_x", "o_FE") if "o_x" in d else ()) + (("o_vx", "o_ed") if "o_vx" in d else ()):
            assert relerr(out[k], d[k]) < 1e-12, (name, k)


@live
def test_reference_scalar_kernels_match_oracle():
    xs = np.concatenate([np.linspace(-2.5, 2.5, 2001), [-2, -1, 0, 1, 2, -1 - 1e-16, 1 + 1e-16]])
    for x in xs:
        assert oracle_py.cubic_bspline(float(x)) == ref_py.cubic_bspline(float(x))        # interpolation.cpp:9-16, bit for bit
        assert oracle_py.dcubic_bspline(float(x)) == ref_py.dcubic_bspline(float(x))      # interpolation.cpp:18-33
    assert ref_py.clamp(0.5, 0.975, 1.0075) == 0.975 and ref_py.clamp(2.0, 0.975, 1.0075) == 1.0075 and ref_py.clamp(1.0, 0.975, 1.0075) == 1.0
    rng = np.random.default_rng(1)
    for _ in range(50):
        A = np.eye(3) + 0.4 * rng.standard_normal((3, 3))
        Qo, Ro = oracle_py.gram_schmidt(A); Qr, Rr = ref_py.gram_schmidt(A)               # geometry.cpp:31-62
        assert np.allclose(Qo, Qr, atol=1e-14) and np.allclose(Ro, Rr, atol=1e-14)
    O = oracle_py.lib(); dp = C.POINTER(C.c_double)
    for kind, par in ((1, [0.3, 0, 0, 0, 0, 0, 0, 0]), (2, [0.9, 0.8, 0.1, 0, 0, 0, 0, 0])):
        P = np.array(par, np.float64)
        for _ in range(300):
            x = rng.random(3); n = np.array([0.0, 0.0, 1.0])
            assert O.orc_ls_phi(C.c_int(kind), P.ctypes.data_as(dp), x.ctypes.data_as(dp)) == ref_py.ls_phi(kind, P, x)     # LevelSet.cpp:8-21
            O.orc_ls_normal(C.c_int(kind), P.ctypes.data_as(dp), x.ctypes.data_as(dp), n.ctypes.data_as(dp))
            assert np.array_equal(n, ref_py.ls_normal(kind, P, x))                                                          # LevelSet.cpp:13-16,23-42


@live
def test_standin_svd_contract():
    """The reference's results in this build rest on MiniEigen's JacobiSVD: hold it to Eigen's contract and to LAPACK."""
    rng = np.random.default_rng(2)
    mats = [np.eye(3), np.zeros((3, 3)), np.diag([2.0, 2.0, 0.5]), np.diag([1.0, -3.0, 2.0]), np.outer([1, 2, 3.0], [0.5, -1, 2.0])]
    mats += [np.eye(3) + s * rng.standard_normal((3, 3)) for s in (1e-8, 1e-3, 0.1, 1.0, 10.0) for _ in range(40)]
    for A in mats:
        U, s, V = ref_py.svd3(A)
        assert np.allclose(U @ np.diag(s) @ V.T, A, atol=1e-13 * max(1.0, np.abs(A).max()))
        assert np.allclose(U.T @ U, np.eye(3), atol=1e-13) and np.allclose(V.T @ V, np.eye(3), atol=1e-13)
        assert (s >= 0).all() and s[0] >= s[1] >= s[2]
        assert np.allclose(s, np.linalg.svd(A, compute_uv=False), atol=1e-13 * max(1.0, np.abs(A).max()))
    for _ in range(100):
        A = rng.standard_normal((2, 2)); U, s, V = ref_py.svd2(A)
        assert np.allclose(U @ np.diag(s) @ V.T, A, atol=1e-13) and np.allclose(U.T @ U, np.eye(2), atol=1e-13) and s[0] >= s[1] >= 0


@live
@pytest.mark.parametrize("material", ["sand", "snow"])
def test_reference_stages_match_oracle_from_identical_state(material):
    """Every stage of the loop body on its own: both sides start each stage from the SAME state (the oracle's), so an error
    in one stage cannot hide behind another.  Fresh seed, not a fixture."""
    from anisotropicelastoplasticity_b200 import scenes as sc
    from oracle.ref_py import Reference
    scene = sc.small_block(material=sc.SAND if material == "sand" else sc.SNOW, res=14, cells=3, seed=77, lo=(0.36, 0.36, 0.22))
    o = Oracle(scene); r = Reference(scene); o.init(); r.init()
    for _ in range(3):                                      # evolve: non-trivial F_E, F_P, q, B
        o.substep(); r.substep()
    dt = 2.5e-4
    o.stage_forces(dt); r.stage_forces(dt)                                           # HS:252-458
    assert relerr(r.grid()["f"], o.grid()["f"]) < 1e-11
    o.stage_grid_update(dt); r.stage_grid_update(dt)                                 # HS:725-737
    assert relerr(mom(**_mv(r)), mom(**_mv(o))) < 1e-11
    assert r.cfl_condition() == pytest.approx(o.cfl_condition(), rel=1e-6)           # RegularGrid.cpp:188-200
    v_before = o.grid()["v"].copy()
    o.stage_collide(); r.stage_collide()                                             # HS:460-511
    go, gr = o.grid(), r.grid()
    assert relerr(mom(gr["m"], gr["v"]), mom(go["m"], go["v"])) < 1e-11 and relerr(mom(gr["m"], gr["vt"]), mom(go["m"], go["vt"])) < 1e-11
    assert (np.abs(go["v"] - v_before).sum(axis=1) > 0).any()                        # the collider did act
    o.stage_g2p(dt); r.stage_g2p(dt)                                                 # HS:739-825, 940-951, 553-681
    po, pr = o.particles(), r.particles()
    for k in ("x", "v", "B", "FE", "FP", "q"):
        assert relerr(pr[k], po[k]) < 1e-10, k
    o.rebuild_weights(); r.rebuild_weights(); o.p2g(False); r.p2g(False)             # HS:18-97, 113-240
    go, gr = o.grid(), r.grid()
    assert relerr(gr["m"], go["m"]) < 1e-12 and relerr(mom(gr["m"], gr["v"]), mom(go["m"], go["v"])) < 1e-11


def _mv(s):
    g = s.grid()
    return dict(m=g["m"], v=g["v"])


@live
def test_reference_sparse_algebra_matches_scipy_transcription():
    """The stand-in's SparseMatrix products against scipy.sparse: the reference run here vs oracle/literal_numpy.py (which
    follows the reference's own sparse-matrix formulation) on a fresh coupled cloth + sand scene."""
    from oracle import literal_numpy as ln
    from oracle.make_golden import golden_scenes
    from oracle.ref_py import Reference
    scene = golden_scenes()["cloth_sand"]; scene.particles.v[:, 0] += 0.3
    r = Reference(scene); l = ln.from_scene(scene); r.init(); l.init()
    for _ in range(2):
        assert r.substep() == pytest.approx(l.substep(), rel=1e-11)
    p = r.particles(); m = r.mesh()
    assert relerr(p["x"], l.ps["x"]) < 1e-12 and relerr(p["FE"], l.ps["FE"]) < 1e-11 and relerr(p["q"], l.ps["q"]) < 1e-10
    assert relerr(m["vx"], l.mesh["vx"]) < 1e-12 and relerr(m["ed"][2], l.mesh["ed3"]) < 1e-11
    assert relerr(r.grid()["f"], l.rg.forces) < 1e-11


@live
@pytest.mark.parametrize("material,dt", [("sand", 4e-4), ("snow", 3e-4)])
def test_reference_bulk_statistics_over_many_substeps(material, dt):
    """200 pinned-dt substeps (the length of BASELINE.json's bulk gate): centre of mass, kinetic energy, plastic volume change
    of the oracle against the reference's own code -- at 1e-7, far inside the 1 % the GPU engine is held to against the oracle
    over the same 200 substeps (tests/test_gpu_parity.py::test_200_substeps_pinned_dt), which closes the chain
    engine ~ oracle ~ reference for the long-run gate."""
    from anisotropicelastoplasticity_b200 import scenes as sc
    from oracle.ref_py import Reference
    scene = sc.small_block(material=sc.SAND if material == "sand" else sc.SNOW, res=16, cells=3, seed=5, lo=(0.4, 0.4, 0.3))
    o = Oracle(scene); r = Reference(scene); o.init(); r.init()
    for _ in range(200):
        for s in (o, r):
            s.stage_forces(dt); s.stage_grid_update(dt); s.stage_collide(); s.stage_g2p(dt); s.rebuild_weights(); s.p2g(False)
    po, pr = o.particles(), r.particles(); mass = scene.particles.m
    so, sr = sc.bulk_stats(po["x"], po["v"], mass, po["FP"]), sc.bulk_stats(pr["x"], pr["v"], mass, pr["FP"])
    for a, b in zip(so, sr):
        assert np.allclose(a, b, rtol=1e-7, atol=1e-12)
    assert relerr(po["x"], pr["x"]) < 1e-8


@live
def test_reference_factories_constants():
    """ParticleSystem::{SnowBall,SandBall,SandBlock,SandCylinder}: material constants and total masses (ParticleSystem.cpp:
    149,173-177,209,233-234,296,321-322,371,395-396); positions are randomly seeded by the reference and only range-checked."""
    n = 500
    x, m, c = ref_py.factory(0, (0.5, 0.5, 0.5), (0, 0, 0), 0.1, 0.0, n)
    assert np.allclose(c, [1.4e5, 0.2, 2.5e-2, 7.5e-3, 0.2]) and m.sum() == pytest.approx(100 * 3.14 * 1e-3)
    assert (np.linalg.norm(x - 0.5, axis=1) <= 0.1 + 1e-12).all()
    x, m, c = ref_py.factory(1, (0.5, 0.5, 0.5), (0, 0, 0), 0.1, 0.0, n)
    assert np.allclose(c[:2], [3.537e5, 0.3]) and m.sum() == pytest.approx(1300 * 3.14 * 1e-3)
    x, m, c = ref_py.factory(2, (0.2, 0.2, 0.1), (0.5, 0.6, 0.7), 0.1, 0.0, n)
    assert m.sum() == pytest.approx(1300 * (0.3 * 0.4 * 0.6 - 0.25 * np.pi * 1e-3))
    assert (x >= [0.2, 0.2, 0.1]).all() and (x <= [0.5, 0.6, 0.7]).all()
    assert (np.linalg.norm(x - np.array([0.2, 0.2, 0.4]), axis=1) >= 0.1).all()          # hole centred on the block's corner edge (ParticleSystem.cpp:251)
    x, m, c = ref_py.factory(3, (-0.5, 0.0, 0.1), (0, 0, 0), 0.25, 0.6, n)
    assert m.sum() == pytest.approx(1300 * np.pi * 0.25 ** 2 * 0.6)
    assert (np.hypot(x[:, 0] + 0.5, x[:, 1]) <= 0.25 + 1e-12).all() and (x[:, 2] >= 0.1).all() and (x[:, 2] <= 0.7).all()


# ------------------------------------------------------------------------------------------------ GPU engine vs the reference
def engine_replay(e, d):
    """The engine's stage entry points (aep_stage_forces / _grid / _g2p, aep_p2g) driven with the time steps the reference
    chose: forces and the grid update with the lagged one (HS:873-877), G2P / advection / plasticity with the new one."""
    dt_prev = float(d["dt0"])
    for dt in map(float, d["dts"]):
        e.stage_forces(dt_prev); e.stage_grid(dt_prev); e.stage_g2p(dt); e.p2g(False)
        dt_prev = dt


@pytest.mark.gpu
@pytest.mark.parametrize("name", REF_GOLDEN)
def test_engine_reproduces_reference_fixture(name):
    """libaep_b200.so against what the REFERENCE'S OWN CODE produced (tests/golden/ref_*.npz), 6 passes of the loop body
    HS:867-988 from an identical state, the reference's time steps replayed.  Tolerances: BASELINE.json's 1e-5 norm-wise
    relative for grid mass, particle x and F; 2e-5 for v, momenta and the APIC matrix B after SIX accumulated fp32 substeps.
    Measured on a B200 (profiles/r1_v11_engine_vs_reference_fixtures.txt): x 4e-9..6e-7, v <= 1.3e-6, F_E <= 5.6e-7,
    F_P <= 6.7e-7, grid mass <= 8.5e-7, momentum <= 1.4e-6, cloth vertex velocity <= 1.1e-6."""
    from anisotropicelastoplasticity_b200.engine import Engine
    d, scene = load_golden("ref_" + name)
    e = Engine(scene); e.init()
    assert e.dt == pytest.approx(float(d["dt0"]), rel=2e-6)                             # HS:860
    g0 = e.grid()
    assert relerr(g0["m"], d["g0_m"]) < 1e-5 and relerr(mom(g0["m"], g0["v"]), mom(d["g0_m"], d["g0_v"])) < 1e-5
    if scene.particles is not None:
        assert relerr(e.particles()["vol"], d["vol_init"]) < 1e-5                     # HS:242-249
    engine_replay(e, d)
    g = e.grid()
    assert relerr(g["m"], d["o_gm"]) < 1e-5 and relerr(mom(g["m"], g["v"]), mom(d["o_gm"], d["o_gv"])) < 2e-5
    if scene.particles is not None:
        p = e.particles()
        for k, tol in (("x", 1e-5), ("v", 2e-5), ("FE", 1e-5), ("FP", 1e-5), ("B", 2e-5)):
            assert relerr(p[k], d["o_" + k]) < tol, k
        assert np.abs(p["q"] - d["o_q"]).max() < 5e-5
    if scene.mesh is not None:
        m = e.mesh()
        for k, tol in (("vx", 1e-5), ("ex", 1e-5), ("vv", 2e-5), ("ev", 2e-5), ("ed", 1e-5)):
            assert relerr(m[k], d["o_" + k]) < tol, k
    assert e.clock()["escaped"] == 0
    e.close()


@pytest.mark.gpu
def test_engine_reproduces_reference_solve_frame():
    """One whole 1/60 s frame as HybridSolver::solve itself ran it (64 substeps, dt between 2e-6 and 1e-3, the last one
    clipped to the frame), the reference's time steps replayed on the engine: particle state and the bulk statistics of
    BASELINE.json's long-run gate (1 %) -- the state itself is held far tighter (measured on a B200 after the 64 substeps:
    x 6.1e-7, v 7.9e-6, F_E 2.3e-6, F_P 7.6e-6; profiles/r1_v11_engine_vs_reference_fixtures.txt)."""
    from anisotropicelastoplasticity_b200.engine import Engine
    from anisotropicelastoplasticity_b200.scenes import bulk_stats
    d, scene = load_golden("ref_solve_frame")
    e = Engine(scene); e.init()
    engine_replay(e, d)
    p = e.particles(); mass = scene.particles.m
    assert relerr(p["x"], d["o_x"]) < 1e-5 and relerr(p["v"], d["o_v"]) < 1e-4 and relerr(p["FE"], d["o_FE"]) < 5e-5 and relerr(p["FP"], d["o_FP"]) < 1e-4
    com, ke, jp = bulk_stats(d["o_x"], d["o_v"], mass, d["o_FP"]); ecom, eke, ejp = bulk_stats(p["x"], p["v"], mass, p["FP"])
    assert np.linalg.norm(ecom - com) < 0.01 * np.linalg.norm(com) and eke == pytest.approx(ke, rel=0.01)
    assert (ejp - 1.0) == pytest.approx(jp - 1.0, rel=0.01, abs=1e-6)
    assert np.allclose(d["obj_xyz"], p["x"], rtol=1e-4)                                # the frame file the reference wrote
    e.close()


# ------------------------------------------------------------------------------------------------ host C++ layer vs the reference
@live
def test_host_objmesh_loader_matches_reference(tmp_path):
    """LagrangianMesh::ObjMesh of libaep_host.so against the reference's loader (LagrangianMesh.cpp:197-352) on the same OBJ
    file (comments, vn/vt lines, positions parsed as float like the reference does): vertex/element masses and volumes, rest
    directions, centroids, Lame constants and tan(friction angle).  Pure host code: no GPU involved."""
    import subprocess
    from test_host_cpp import DRIVER, _build, read_blob
    _build()
    rng = np.random.default_rng(5); n = 6
    gx, gy = np.meshgrid(np.linspace(0.1, 0.9, n), np.linspace(0.2, 0.7, n), indexing="ij")
    V = np.stack([gx.ravel(), gy.ravel(), 0.6 + 0.02 * rng.standard_normal(n * n)], axis=1)
    F = []
    for i in range(n - 1):
        for j in range(n - 1):
            a = i * n + j; F += [(a, a + n, a + n + 1), (a, a + n + 1, a + 1)]
    obj = tmp_path / "cloth.obj"
    with open(obj, "w") as f:
        f.write("# cloth\no sheet\n")
        for v in V:
            f.write(f"v {v[0]:.7f} {v[1]:.7f} {v[2]:.7f}\n")
        f.write("vt 0.5 0.5\nvn 0 0 1\n")
        for t in F:
            f.write(f"f {t[0] + 1} {t[1] + 1} {t[2] + 1}\n")
    par = (2e3, 0.04, 200.0, 0.3, 12.0, 4e4, 25.0)                         # main.cpp:77-78 with shear and friction switched on
    ref = ref_py.obj_mesh(str(obj), *par)
    out = tmp_path / "mesh.bin"
    subprocess.check_call([DRIVER, "objmesh", str(obj), str(out)] + [repr(float(p)) for p in par])
    got = read_blob(str(out)); nv, nf = n * n, len(F)
    cm = lambda a, k: a.reshape(3, k).T
    assert np.array_equal(cm(got["faces"], nf).astype(int), ref["faces"]) and np.array_equal(ref["faces"], np.array(F))
    assert np.array_equal(cm(got["vx"], nv), ref["vx"])                     # both parse through float
    for k in ("vm", "vvol", "em", "evol"):
        assert relerr(got[k], ref[k]) < 1e-14, k
    for a, k in enumerate(("D1", "D2", "D3")):
        assert relerr(cm(got[k], nf), ref["eD"][a]) < 1e-14, k
    assert np.allclose(got["consts"], [ref["mu"], ref["lam"], ref["fric"]], rtol=1e-15)
    assert ref["fric"] == pytest.approx(np.tan(np.deg2rad(25.0)), rel=1e-15)   # LagrangianMesh.cpp:351


@live
def test_pinned_vertex_block_wraps_at_domain_face_like_the_reference():
    """Quirk (HS:513-550): around every stencil node of a pinned cloth vertex the reference zeroes a 3x3x3 node block, checking
    only the FLAT index (HS:538-539), so at a domain face i = -1 / i = nx wrap into the neighbouring grid row.  A sheet whose
    pinned corner sits one cell from the x = 0 face, with every node moving: the zeroed node set must be the reference's."""
    from anisotropicelastoplasticity_b200 import scenes as sc
    from oracle.ref_py import Reference
    n = 5; res = 12; h = 1.0 / res
    mesh = sc.make_cloth(n, n, (1.2 * h, 3.3 * h, 6.4 * h), (0.9 * h, 0, 0), (0, 0.9 * h, 0), fixed_ids=(0,))
    mesh.vv[:] = (0.3, -0.2, 0.5)
    g = sc.GridSpec(np.zeros(3), np.ones(3), np.array([res] * 3))
    scene = sc.Scene("pin_at_face", g, sc.SAND, None, mesh, sc.LevelSetSpec())
    o = Oracle(scene); r = Reference(scene); o.init(); r.init()
    ng = g.n_nodes; rng = np.random.default_rng(3)
    m_all = np.full(ng, 0.5); v_all = 1.0 + rng.random((ng, 3))                     # every node of the grid carries mass and moves
    for s in (o, r):
        s.set_grid(m=m_all, v=v_all); s.stage_collide()
    go, gr = o.grid(), r.grid()
    zo = np.abs(go["v"]).sum(axis=1) == 0; zr = np.abs(gr["v"]).sum(axis=1) == 0
    assert zr.sum() > 27 and np.array_equal(zo, zr)
    assert np.array_equal(np.abs(go["vt"]).sum(axis=1) == 0, np.abs(gr["vt"]).sum(axis=1) == 0)     # the pre-friction copy too (HS:542)
    i = np.nonzero(zr)[0] % res
    assert (i == 0).any() and (i == res - 1).any(), "the block did not wrap: move the pinned vertex closer to the face"   # i = -1 landed in the previous row
    assert np.array_equal(go["v"], gr["v"]) and np.array_equal(go["vt"], gr["vt"])


# ------------------------------------------------------------------------------------------------ BASELINE configs[0], exact size
def _c1_exact():
    """BASELINE.json configs[0]: the sand-block drop, 104 361 particles on a 64^3 grid -- "runnable on the reference CPU path",
    so it is run there.  The state is strained and moving so that the SVDs, the Drucker-Prager projection and the APIC terms all work."""
    from anisotropicelastoplasticity_b200 import scenes as sc
    s = sc.c1_sand_block(res=64)
    sc.perturb_state(s.particles, np.random.default_rng(3), strain=1e-2, vel=0.3, affine=1.0)
    return s


def _reference_run(scene, nsteps):
    """dt0, dts and final state of the reference's own code; same keys as the fixtures."""
    from oracle.make_ref_golden import run_reference
    return run_reference(scene, nsteps)


@live
def test_c1_exact_oracle_vs_reference():
    scene = _c1_exact(); d = _reference_run(scene, 2)
    o = Oracle(scene, threads=0); o.init()
    assert o.dt == pytest.approx(float(d["dt0"]), rel=1e-12) and relerr(o.particles()["vol"], d["vol_init"]) < 1e-13
    replay(o, scene, d["dt0"], d["dts"], rule_rel=1e-6)
    compare_states(o, scene, d)


@live
@pytest.mark.parametrize("material", ["sand", "snow"])
def test_degenerate_deformation_gradients_match_reference(material):
    """Particles whose F_E is special: identity (all singular values equal), two equal singular values, a pure rotation, a
    reflection (det < 0: one singular vector flips sign), strong compression / stretch, nearly singular.  The SVD is not unique
    there; every use of it in the reference (stress HS:314-339, return mapping HS:626-677) must not care.  Particles exactly on
    a grid node and on a cell face exercise the `N > 0` filter of HS:60 and the truncation of HS:34-36."""
    from anisotropicelastoplasticity_b200 import scenes as sc
    from oracle.ref_py import Reference
    from random_scenes import degenerate_scene
    scene, n_special = degenerate_scene(sc.SAND if material == "sand" else sc.SNOW)
    p = scene.particles
    o = Oracle(scene); r = Reference(scene); o.init(); r.init()
    assert relerr(r.grid()["m"], o.grid()["m"]) < 1e-13 and relerr(r.particles()["vol"], o.particles()["vol"]) < 1e-13
    dt = 2e-4
    for s in (o, r):
        s.stage_forces(dt)
    assert relerr(r.grid()["f"], o.grid()["f"]) < 1e-10
    for s in (o, r):
        s.stage_grid_update(dt); s.stage_collide(); s.stage_g2p(dt)
    po, pr = o.particles(), r.particles()
    for k in ("x", "v", "B", "q"):
        assert relerr(pr[k], po[k]) < 1e-10, k
    # F_E F_P (the total deformation) is what the SVD's freedom cannot touch; F_E and F_P individually agree too unless the
    # reflection case puts the sign in a different factor
    tot = lambda P: np.einsum("nij,njk->nik", P["FE"], P["FP"])
    assert relerr(tot(pr), tot(po)) < 1e-10
    keep = np.ones(p.n, bool); keep[3] = False
    assert relerr(pr["FE"][keep], po["FE"][keep]) < 1e-9 and relerr(pr["FP"][keep], po["FP"][keep]) < 1e-9
    assert np.isfinite(pr["FE"]).all() and np.isfinite(po["FE"]).all()


@live
def test_reference_solve_frame_numbering(tmp_path):
    """`while (t <= maxt)` with t += 1/60 per finished frame (HS:867,883): maxt = 1.5/60 writes particle_0.obj and particle_1.obj and
    stops -- the convention HybridSolver::solve of libaep_host.so reproduces (tests/test_host_cpp.py, maxt = n/60 - 1/120 -> n frames)."""
    from anisotropicelastoplasticity_b200 import scenes as sc
    from oracle.ref_py import Reference
    scene = sc.small_block(material=sc.SAND, res=12, cells=2, seed=61, lo=(0.34, 0.34, 0.34))
    r = Reference(scene); r.solve(1.5 / 60.0, str(tmp_path))
    assert sorted(os.listdir(tmp_path / "particle")) == ["particle_0.obj", "particle_1.obj"]
    assert os.listdir(tmp_path / "mesh") == []                                    # the directory is made (HS:858), no mesh bound
    lines = open(tmp_path / "particle" / "particle_1.obj").read().splitlines()
    assert len(lines) == scene.particles.n and all(ln.startswith("v ") and len(ln.split()) == 4 for ln in lines)
    assert np.allclose(np.array([[float(t) for t in ln.split()[1:]] for ln in lines]), r.particles()["x"], rtol=2e-5)


def _replay_pair(o, r, nsteps=2):
    """nsteps of the reference free-running, the oracle replaying its time steps (forces / grid update with the lagged one)."""
    dts = [r.substep() for _ in range(nsteps)]
    dt_lag = float(o.dt)
    for dt in dts:
        o.stage_forces(dt_lag); o.stage_grid_update(dt_lag); o.stage_collide(); o.stage_g2p(float(dt)); o.rebuild_weights(); o.p2g(False); dt_lag = float(dt)


@live
def test_random_scenes_oracle_vs_reference():
    """40 seeded random particle scenes (tests/random_scenes.py): init + 2 substeps, time steps replayed.  A sweep for disagreements
    in corners no hand-made scene visits."""
    from oracle.ref_py import Reference
    from random_scenes import random_particle_scene
    worst = 0.0
    for seed in range(40):
        scene = random_particle_scene(seed)
        o = Oracle(scene); r = Reference(scene); o.init(); r.init()
        assert o.dt == pytest.approx(r.dt, rel=1e-9), seed
        assert relerr(o.particles()["vol"], r.particles()["vol"]) < 1e-12, seed
        _replay_pair(o, r)
        po, pr = o.particles(), r.particles(); go, gr = o.grid(), r.grid()
        errs = [relerr(po[k], pr[k]) for k in ("x", "v", "B", "FE", "FP")] + [relerr(go["m"], gr["m"]), relerr(mom(go["m"], go["v"]), mom(gr["m"], gr["v"])), relerr(go["f"], gr["f"])]
        errs.append(float(np.abs(po["q"] - pr["q"]).max()))
        assert max(errs) < 1e-9, (seed, errs)
        worst = max(worst, max(errs))
    print("worst relative error over 40 random scenes:", worst)


@live
def test_random_cloth_scenes_oracle_vs_reference():
    """24 seeded random cloth scenes (tests/random_scenes.py), with and without sand: init + 2 substeps, time steps replayed."""
    from oracle.ref_py import Reference
    from random_scenes import random_cloth_scene
    worst = 0.0
    for seed in range(24):
        scene = random_cloth_scene(seed)
        o = Oracle(scene); r = Reference(scene); o.init(); r.init()
        assert o.dt == pytest.approx(r.dt, rel=1e-9), seed
        _replay_pair(o, r)
        mo, mr = o.mesh(), r.mesh(); go, gr = o.grid(), r.grid()
        errs = [relerr(mo[k], mr[k]) for k in ("vx", "vv", "vB", "ex", "ev", "eB", "ed")] + [relerr(go["m"], gr["m"]), relerr(mom(go["m"], go["v"]), mom(gr["m"], gr["v"])), relerr(go["f"], gr["f"])]
        if scene.particles is not None:
            po, pr = o.particles(), r.particles()
            errs += [relerr(po[k], pr[k]) for k in ("x", "v", "FE", "FP")]
        assert max(errs) < 1e-9, (seed, errs)
        worst = max(worst, max(errs))
    print("worst relative error over 24 random cloth scenes:", worst)
