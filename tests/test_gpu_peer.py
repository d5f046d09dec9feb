"""The peer-memory slab exchange (aep_comm_*: halo planes, migrating particles, max|v| stored into the neighbour's memory by the kernels
of the fused substep, epoch flags, device-side particle counts) on ONE GPU: several slab contexts of this process connected by plain
pointers -- the same kernels and protocol as one process per GPU over CUDA IPC (tools/peer_parity.py runs that under torchrun).
Gate: the union of the slabs equals the whole-domain context and the oracle."""
import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu


def _scene(res=32, vy=3.0):
    """C1-like sand block, perturbed and drifting along +y so that particles cross slab boundaries; ordered by y (contiguous ids)."""
    from anisotropicelastoplasticity_b200 import scenes as sc
    s = sc.c1_sand_block(res=res)
    sc.perturb_state(s.particles, np.random.default_rng(5), strain=5e-3, vel=0.3, affine=1.0)
    s.particles.v[:, 1] += vy
    p = s.particles; order = np.argsort(p.x[:, 1], kind="stable")
    for k in ("x", "v", "B", "FE", "FP", "m", "vol", "q"):
        setattr(p, k, getattr(p, k)[order])
    return s


def _peer_group(scene, nslabs, axis=1, **kw):
    from anisotropicelastoplasticity_b200 import capi
    from anisotropicelastoplasticity_b200.distributed import PeerSlabGroup, SlabPlan, make_gpu_slab_engine
    cells = np.floor((scene.particles.x[:, axis] - scene.grid.mn[axis]) / scene.grid.h[axis]).astype(np.int64)
    plan = SlabPlan.balanced(cells, int(scene.grid.res[axis]), nslabs, axis=axis)
    engs = []; n0 = []
    for r in range(nslabs):
        eng, local, idx = make_gpu_slab_engine(scene, plan, r, device=0, **kw)
        capi.check(eng.L.aep_set_particle_id_base(eng.h, 0), eng.h)
        # global ids = position in the scene arrays: upload with an explicit id base per contiguous range
        assert (np.diff(idx) == 1).all()
        capi.check(eng.L.aep_set_particle_id_base(eng.h, int(idx[0])), eng.h)
        eng.upload_particles(local)
        engs.append(eng); n0.append(len(idx))
    return PeerSlabGroup(engs, migrate_capacity=4096), engs, n0, plan


@pytest.mark.parametrize("nslabs", [2, 3])
def test_peer_slabs_pinned_dt_match_single_context_and_oracle(nslabs):
    from anisotropicelastoplasticity_b200.engine import Engine
    from oracle.oracle_py import Oracle
    scene = _scene(); dt = float(np.float32(2e-4)); nsteps = 12
    grp, engs, n0, plan = _peer_group(scene, nslabs)
    grp.init()
    for e in engs:
        e.set_fixed_dt(dt)
    grp.run(nsteps)
    got = grp.gather_particles()
    assert (got["ids"] == np.arange(scene.particles.n)).all()                                  # nobody lost or duplicated
    assert any(e.n_particles != n for e, n in zip(engs, n0)), "no particle migrated: test is vacuous"
    assert all(e.clock()["escaped"] == 0 for e in engs)
    whole = Engine(scene); whole.init(); whole.set_fixed_dt(dt); whole.run(nsteps); pw = whole.particles()
    o = Oracle(scene, threads=0); o.init()
    for _ in range(nsteps):
        o.stage_forces(dt); o.stage_grid_update(dt); o.stage_collide(); o.stage_g2p(dt); o.rebuild_weights(); o.p2g(False)
    po = o.particles()
    for k, tol in (("x", 1e-6), ("v", 2e-5), ("FE", 1e-5), ("FP", 1e-5), ("q", 1e-4)):
        assert relerr(got[k], pw[k]) < tol, ("vs whole", k, relerr(got[k], pw[k]))
    for k, tol in (("x", 1e-5), ("v", 1e-4), ("FE", 1e-5), ("FP", 1e-5)):
        assert relerr(got[k], po[k]) < tol, ("vs oracle", k, relerr(got[k], po[k]))


def test_peer_slabs_adaptive_dt_is_global():
    """The reference dt rule needs the GLOBAL max|v_i| (HybridSolver.cpp:878): every slab must hold the dt of the whole-domain context."""
    from anisotropicelastoplasticity_b200.engine import Engine
    scene = _scene()
    grp, engs, n0, plan = _peer_group(scene, 3)
    whole = Engine(scene); whole.init()
    grp.init()
    for step in range(6):
        for e in engs:
            assert e.clock()["dt"] == pytest.approx(whole.clock()["dt"], rel=2e-4), step
        grp.substep(); whole.substep()
    assert sum(e.n_particles for e in engs) == scene.particles.n
    ms = sum(e.stats()["mass"] for e in engs)
    assert ms == pytest.approx(whole.stats()["mass"], rel=1e-6)


def test_peer_slabs_z_axis_and_resort():
    """z-slabs (the other supported axis) over enough substeps that re-sorts and compactions of dead slots happen."""
    from anisotropicelastoplasticity_b200.engine import Engine
    from anisotropicelastoplasticity_b200 import scenes as sc
    s = sc.c1_sand_block(res=32)
    sc.perturb_state(s.particles, np.random.default_rng(9), strain=5e-3, vel=0.3, affine=1.0)
    s.particles.v[:, 2] -= 2.0
    p = s.particles; order = np.argsort(p.x[:, 2], kind="stable")
    for k in ("x", "v", "B", "FE", "FP", "m", "vol", "q"):
        setattr(p, k, getattr(p, k)[order])
    dt = float(np.float32(2e-4)); nsteps = 40
    grp, engs, n0, plan = _peer_group(s, 2, axis=2, sort_every=7)
    grp.init()
    for e in engs:
        e.set_fixed_dt(dt)
    grp.run(nsteps)
    got = grp.gather_particles()
    assert (got["ids"] == np.arange(s.particles.n)).all()
    assert sum(e.counters()["sorts"] for e in engs) >= 2 * (nsteps // 7)
    whole = Engine(s, sort_every=7); whole.init(); whole.set_fixed_dt(dt); whole.run(nsteps); pw = whole.particles()
    for k, tol in (("x", 2e-6), ("v", 1e-4), ("FE", 3e-5)):
        assert relerr(got[k], pw[k]) < tol, (k, relerr(got[k], pw[k]))


def _coupling_scene(res=48, cloth_n=40):
    """configs[3] in small: sand block over a cloth pinned at two corners; the sheet spans the whole y range (every slab owns a strip
    of it), the sand is lowered onto it and everything drifts along +y so that particles AND cloth points change their owner."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_zzx_configs_at_size import deform_cloth
    from anisotropicelastoplasticity_b200 import scenes as sc
    s = sc.c4_coupling(res=res, cloth_n=cloth_n)
    rng = np.random.default_rng(17)
    s.particles.x[:, 2] -= 0.05 - 1.5 / res
    sc.perturb_state(s.particles, rng, strain=5e-3, vel=0.2, affine=0.5)
    s.particles.v[:, 1] += 2.5; s.particles.v[:, 2] -= 1.0
    deform_cloth(s.mesh, rng, amp=0.01, vel=0.3)
    s.mesh.vv[:, 1] += 5.0; s.mesh.ev[:, 1] += 5.0                                 # a row of vertices crosses the slab boundary within 20 substeps
    p = s.particles; order = np.argsort(p.x[:, 1], kind="stable")
    for k in ("x", "v", "B", "FE", "FP", "m", "vol", "q"):
        setattr(p, k, getattr(p, k)[order])
    return s


@pytest.mark.parametrize("nslabs", [2, 4])
def test_peer_slabs_cloth_sand_coupling_matches_single_context(nslabs):
    """Cloth in slab contexts (SURVEY 8e "Cloth", BASELINE configs[3] "1/2/4 B200"): every rank holds the whole mesh state, transfers
    the vertices / elements whose cell lies in its slab (HS:121-125, 137-141, 378, 444-454), computes the in-plane forces redundantly,
    and pushes the points it advanced into every other rank's copy.  Gate: the slabs' union equals the whole-domain context."""
    from anisotropicelastoplasticity_b200.engine import Engine
    scene = _coupling_scene()
    dt = float(np.float32(1e-4)); nsteps = 20
    ycell = np.floor(scene.mesh.vx[:, 1] * scene.grid.res[1]).astype(int)
    grp, engs, n0, plan = _peer_group(scene, nslabs)
    owners0 = plan.owner_of_cells(ycell)
    assert len(set(owners0.tolist())) == nslabs                                    # the sheet lies in every slab
    grp.init()
    for e in engs:
        e.set_fixed_dt(dt)
    grp.run(nsteps)
    got = grp.gather_particles()
    assert (got["ids"] == np.arange(scene.particles.n)).all()
    assert any(e.n_particles != n for e, n in zip(engs, n0)), "no particle migrated: test is vacuous"
    whole = Engine(scene); whole.init(); whole.set_fixed_dt(dt); whole.run(nsteps); pw = whole.particles(); mw = whole.mesh()
    ycell1 = np.floor(mw["vx"][:, 1] * scene.grid.res[1]).astype(int)
    assert (plan.owner_of_cells(ycell1) != owners0).any(), "no cloth vertex changed its owner: test is vacuous"
    for k, tol in (("x", 2e-6), ("v", 5e-5), ("FE", 2e-5), ("FP", 2e-5)):
        assert relerr(got[k], pw[k]) < tol, ("particles vs whole", k, relerr(got[k], pw[k]))
    for r, e in enumerate(engs):                                                     # every rank's copy of the mesh is complete and current
        me = e.mesh()
        for k, tol in (("vx", 2e-6), ("vv", 1e-4), ("ex", 2e-6), ("ev", 1e-4), ("ed", 5e-5)):
            assert relerr(me[k], mw[k]) < tol, ("mesh vs whole", r, k, relerr(me[k], mw[k]))
    assert all(e.clock()["escaped"] == 0 for e in engs)
