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
