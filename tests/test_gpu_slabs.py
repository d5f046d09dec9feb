"""Decomposition invariance on the GPU: several slab contexts (C-ABI halo / migration entry points) against one whole-domain
context and against the oracle.  Runs on ONE GPU through LocalSlabGroup (direct buffer swaps); the NCCL path over
torch.distributed is exercised by `bench.py --gpus N` and tests/test_distributed_cpu.py (same driver code, gloo)."""
import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu


def _scene(res=32):
    """C1-like sand block, perturbed and drifting along +y; particles ordered by y so that slabs own contiguous id ranges."""
    from anisotropicelastoplasticity_b200 import scenes as sc
    s = sc.c1_sand_block(res=res)
    sc.perturb_state(s.particles, np.random.default_rng(5), strain=5e-3, vel=0.3, affine=1.0)
    s.particles.v[:, 1] += 3.0
    p = s.particles; order = np.argsort(p.x[:, 1], kind="stable")
    for k in ("x", "v", "B", "FE", "FP", "m", "vol", "q"):
        setattr(p, k, getattr(p, k)[order])
    return s


def _group(scene, nslabs, overlapped=None):
    from anisotropicelastoplasticity_b200 import capi
    from anisotropicelastoplasticity_b200.distributed import GpuSlabBackend, LocalSlabGroup, SlabPlan, make_gpu_slab_engine
    cells = np.floor(scene.particles.x[:, 1] / scene.grid.h[1]).astype(np.int64)
    plan = SlabPlan.balanced(cells, int(scene.grid.res[1]), nslabs)
    backends = []; n0 = []
    for r in range(nslabs):
        eng, local, idx = make_gpu_slab_engine(scene, plan, r, device=0)
        assert (np.diff(idx) == 1).all()
        capi.check(eng.L.aep_set_particle_id_base(eng.h, int(idx[0])), eng.h)      # global ids = position in the scene arrays
        eng.upload_particles(local)
        backends.append(GpuSlabBackend(eng, migrate_capacity=4096)); n0.append(len(idx))
    return LocalSlabGroup(backends, plan, overlapped=overlapped), backends, n0


@pytest.mark.parametrize("nslabs,overlapped", [(2, True), (3, True), (3, False)])
def test_slabs_pinned_dt_match_single_context_and_oracle(nslabs, overlapped):
    """12 substeps at a pinned dt (well-conditioned, see test_gpu_parity.test_200_substeps_pinned_dt): the union of the slabs,
    gathered by global id, equals the whole-domain context (same fp32 arithmetic, different summation order) and the oracle."""
    from anisotropicelastoplasticity_b200.engine import Engine
    from oracle.oracle_py import Oracle
    scene = _scene(); dt = float(np.float32(2e-4)); nsteps = 12
    grp, backends, n0 = _group(scene, nslabs, overlapped)      # overlapped: leaver lists from G2P + split P2G; else full-scan extract
    grp.init()
    for b in backends:
        b.e.set_fixed_dt(dt)
    grp.run(nsteps)
    got = grp.gather_particles()
    assert (got["ids"] == np.arange(scene.particles.n)).all()                                  # nobody lost or duplicated
    assert any(b.e.n_particles != n for b, n in zip(backends, n0)), "no particle migrated: test is vacuous"
    assert all(b.e.clock()["escaped"] == 0 for b in backends)
    whole = Engine(scene); whole.init(); whole.set_fixed_dt(dt); whole.run(nsteps); pw = whole.particles()
    o = Oracle(scene, threads=0); o.init()
    for _ in range(nsteps):
        o.stage_forces(dt); o.stage_grid_update(dt); o.stage_collide(); o.stage_g2p(dt); o.rebuild_weights(); o.p2g(False)
    po = o.particles()
    for k, tol in (("x", 1e-6), ("v", 2e-5), ("FE", 1e-5), ("FP", 1e-5), ("q", 1e-4)):
        assert relerr(got[k], pw[k]) < tol, ("vs whole", k, relerr(got[k], pw[k]))
    for k, tol in (("x", 1e-5), ("v", 1e-4), ("FE", 1e-5), ("FP", 1e-5)):
        assert relerr(got[k], po[k]) < tol, ("vs oracle", k, relerr(got[k], po[k]))


def test_slabs_adaptive_dt_allreduce():
    """The reference dt rule needs the GLOBAL max|v_i| (HybridSolver.cpp:878): after the all-reduce every slab holds the dt of
    the whole-domain context.  (Only dt and conservation are compared: with the adaptive rule the reference is chaotic.)"""
    from anisotropicelastoplasticity_b200.engine import Engine
    scene = _scene()
    grp, backends, n0 = _group(scene, 3)
    whole = Engine(scene); whole.init()
    grp.init()
    for step in range(6):
        for b in backends:
            assert b.e.clock()["dt"] == pytest.approx(whole.clock()["dt"], rel=2e-4), step
        grp.substep(); whole.substep()
    assert sum(b.e.n_particles for b in backends) == scene.particles.n
    ms = sum(b.e.stats()["mass"] for b in backends)
    assert ms == pytest.approx(whole.stats()["mass"], rel=1e-6)
