"""Checkpoint / restart (the reference has none): aep_resume + aep_set_clock behind Engine.checkpoint /
Engine.resume and HybridSolver::saveCheckpoint / resume.  A run that is saved after n substeps and continued in a NEW context must
land where the uninterrupted run lands.  What can differ: the re-sort of the re-uploaded particles changes the order of the
fp32 atomic sums (~1e-7 per substep), nothing else -- the fp64 boundary arrays carry the fp32 device state exactly.
(The file runs last on purpose: it exercises the newest entry points.)"""
import os

import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu


def _scene(kind):
    from anisotropicelastoplasticity_b200 import scenes as sc
    if kind == "cloth_sand":
        from conftest import load_golden
        return load_golden("cloth_sand")[1]
    return sc.small_block(material=sc.SAND if kind == "sand" else sc.SNOW, res=16, cells=3, seed=41)


@pytest.mark.parametrize("kind", ["sand", "snow", "cloth_sand"])
def test_engine_resume_continues_the_run(kind, tmp_path):
    from anisotropicelastoplasticity_b200.engine import Engine
    scene = _scene(kind); dt = 2.5e-4; n = 8
    a = Engine(scene); a.init(); a.set_fixed_dt(dt); a.run(n)
    ck = a.checkpoint(); ca = a.clock()
    path = str(tmp_path / "ck.npz"); np.savez(path, **ck)              # through a file, like a real restart
    a.run(n); pa = a.particles(); ma = a.mesh() if scene.mesh is not None else None; ca2 = a.clock(); a.close()
    b = Engine.resume(scene, dict(np.load(path)))
    cb = b.clock()
    for k in ("dt", "t", "inner_t", "frame", "substeps"):
        assert cb[k] == ca[k], k                                         # the clock is back exactly
    g = b.grid(); assert g["m"].sum() == pytest.approx(scene_mass(scene), rel=1e-5)      # aep_resume ran the P2G
    assert relerr(b.particles()["vol"], ck["p_vol"]) == 0.0            # volumes are state, not recomputed (HS:242-249 runs once)
    b.set_fixed_dt(dt); b.run(n); pb = b.particles(); cb2 = b.clock()
    assert cb2["substeps"] == ca2["substeps"] == 2 * n and cb2["frame"] == ca2["frame"]
    assert cb2["t"] + cb2["inner_t"] == pytest.approx(ca2["t"] + ca2["inner_t"], rel=1e-12)
    for k, tol in (("x", 1e-6), ("v", 1e-4), ("FE", 1e-5), ("FP", 1e-5), ("B", 1e-3)):
        assert relerr(pb[k], pa[k]) < tol, k
    if ma is not None:
        mb = b.mesh()
        for k, tol in (("vx", 1e-6), ("vv", 1e-4), ("ed", 1e-5)):
            assert relerr(mb[k], ma[k]) < tol, k
    b.close()


def scene_mass(scene):
    m = 0.0
    if scene.particles is not None:
        m += float(scene.particles.m.sum())
    if scene.mesh is not None:
        m += float(scene.mesh.vm.sum() + scene.mesh.em.sum())
    return m


def test_set_clock_rejects_nonsense():
    from anisotropicelastoplasticity_b200 import capi
    from anisotropicelastoplasticity_b200.engine import Engine
    e = Engine(_scene("sand")); e.init()
    assert e.L.aep_set_clock(e.h, 0.0, 0.0, 0.0, 0, 0) == -1 and e.L.aep_set_clock(e.h, 1e-4, 0.0, -1.0, 0, 0) == -1
    assert e.L.aep_set_clock(e.h, 1e-4, 0.5, 0.001, 30, 777) == 0
    c = e.clock()
    assert c["dt"] == float(np.float32(1e-4)) and c["t"] == 0.5 and c["inner_t"] == 0.001 and c["frame"] == 30 and c["substeps"] == 777
    e.close()


def test_hybrid_solver_checkpoint_resume(tmp_path):
    """The C++ host classes: begin, n substeps, saveCheckpoint, n substeps | new process: resume from the file, n substeps."""
    from anisotropicelastoplasticity_b200 import scenes as sc
    from test_host_cpp import run_driver
    scene = sc.small_block(material=sc.SAND, res=16, cells=3, seed=42)
    outdir = str(tmp_path / "ck"); os.makedirs(outdir)
    extra = {"fixed_dt": np.array([2.5e-4])}
    a = run_driver(tmp_path, scene, "ckpt_save", 6, tag="save", outdir=outdir, extra=extra)
    assert os.path.getsize(os.path.join(outdir, "state.ckpt")) > 36 * 8 * scene.particles.n
    b = run_driver(tmp_path, scene, "ckpt_resume", 6, tag="resume", outdir=outdir, extra=extra)
    assert a["info"][3] == 12 and b["info"][3] == 12 and a["info"][1] == pytest.approx(b["info"][1], rel=1e-12)      # substeps, simulated time
    for k, tol in (("x", 1e-6), ("v", 1e-4), ("FE", 1e-5), ("FP", 1e-5)):
        assert relerr(b[k], a[k]) < tol, k
    assert relerr(b["vol"], a["vol"]) == 0.0 and relerr(b["grid_m"], a["grid_m"]) < 1e-5
