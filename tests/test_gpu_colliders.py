"""SURVEY 8(f) rank 3 -- the level-set library beyond the reference's static stick-or-slide colliders, each with a known-answer test on
the grid the engine hands back through the C ABI:

  * moving colliders (aep_set_collider_motion): the projection HS:484-502 works on the velocity RELATIVE to the collider; a node that
    sticks moves with the collider.  The reference's colliders are static (HS:484 "for static object"): velocity 0 reproduces it bit for bit.
  * opt-in Coulomb friction (aep_config.coulomb_friction): what HS:500-501 set out to do (the reference's statement there is a no-op):
    the tangential velocity shrinks by mu |v_n| instead of the all-or-nothing stick test.  Off by default: parity with the reference.
  * opt-in mass floor for the dt rule (aep_config.vmax_min_mass_fraction; NOT the reference rule RegularGrid.cpp:188-200): nodes lighter
    than a fraction of one particle's mass do not enter max|v|.

Known answers come from a twin engine WITHOUT a collider run through the same stages: its post-update grid velocity is the collider's
input (HS:463), to which the formulas are applied in numpy."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _scene(levelset=True):
    """sand block dropped INTO a ground plane at z0 (several node layers of the block lie below it), moving sideways and down"""
    from anisotropicelastoplasticity_b200 import scenes as sc
    s = sc.small_block(material=sc.SAND, res=24, cells=5, seed=3, perturb=False, levelset=False, lo=(0.3, 0.3, 0.25))
    rng = np.random.default_rng(4)
    sc.perturb_state(s.particles, rng, strain=2e-3, vel=0.0, affine=0.0)
    ramp = ((s.particles.x[:, 0] - 0.3) / (5.0 / 24.0))[:, None]                           # tangential speed 0 .. full across the block:
    s.particles.v = np.array([0.6, -0.25, 0.0]) * ramp + np.array([0.0, 0.0, -0.8]) + 0.05 * rng.standard_normal(s.particles.v.shape)   # both sticking and sliding nodes
    if levelset:
        s.levelset = sc.LevelSetSpec(sc.LS_GROUND, np.array([0.31, 0, 0, 0, 0, 0, 0, 0.0]))
    return s


def _after_grid_update(scene, dt, **kw):
    from anisotropicelastoplasticity_b200.engine import Engine
    motion = kw.pop("motion", None)
    e = Engine(scene, **kw)
    if motion is not None:
        e.set_collider_motion(motion)
    e.init(); e.stage_forces(dt); e.stage_grid(dt)
    g = e.grid(); clk = e.clock()
    return e, g, clk


def _node_z(scene):
    g = scene.grid; n = np.arange(g.n_nodes)
    return g.mn[2] + (n // (g.res[0] * g.res[1])) * g.h[2]


MU = 0.2          # aep_config.collider_friction (HS:498)


@pytest.mark.parametrize("vc", [(0.0, 0.0, 0.0), (0.3, -0.1, 0.5), (0.1, 0.2, -0.3)])
def test_moving_ground_projects_the_relative_velocity(vc):
    dt = 2e-4; vc = np.asarray(vc, float)
    free = _after_grid_update(_scene(levelset=False), dt)[1]                                        # HS:463: the collider's input
    e, g, clk = _after_grid_update(_scene(), dt, motion=vc)
    z = _node_z(_scene()); inside = (z - 0.31 <= 0.0) & (free["m"] > 0)                             # t = 0: the plane has not moved yet
    assert inside.sum() > 200
    n = np.array([0.0, 0.0, 1.0])
    rel = free["v"] - vc; vn = rel @ n
    hit = inside & (vn < 0)
    assert hit.sum() > 100
    rel_t = rel - vn[:, None] * n
    vt_expected = np.where(hit[:, None], rel_t + vc, free["v"])                                     # v~ (HS:490-492), lab frame
    stick = hit & (np.linalg.norm(rel_t, axis=1) < -MU * vn)                                        # HS:494-502
    v_expected = np.where(stick[:, None], vc, vt_expected)                                          # a sticking node moves with the collider
    act = free["m"] > 1e-12 * free["m"].max()
    scale = np.abs(free["v"][act]).max()
    assert np.abs(g["vt"][act] - vt_expected[act]).max() < 2e-6 * scale
    assert np.abs(g["v"][act] - v_expected[act]).max() < 2e-6 * scale
    assert (hit & ~stick).sum() > 10 and (stick.sum() > 10 or vc[2] < 0)                            # both branches were taken (a receding plane: everything slides)
    if not vc.any():                                                                                # velocity 0 = the reference's static collider (same code path;
        s = _after_grid_update(_scene(), dt)[1]                                                     # two runs differ by the order of the float atomics only)
        assert np.abs(s["v"] - g["v"]).max() < 1e-6 * scale and np.abs(s["vt"] - g["vt"]).max() < 1e-6 * scale
    e.close()


def test_moving_collider_carries_particles_and_advances_in_time():
    """A plane rising at 1 m/s under a resting block: after n substeps the plane stands at z0 + w t, the block's lowest particles move
    up with it, and particles G2P'd from sticking nodes take the collider's velocity (not zero)."""
    from anisotropicelastoplasticity_b200.engine import Engine
    scene = _scene(); scene.particles.v[:] = 0.0
    w = 1.0; dt = float(np.float32(2e-4)); nsteps = 60
    e = Engine(scene); e.set_collider_motion((0.0, 0.0, w)); e.init(); e.set_fixed_dt(dt); e.run(nsteps)
    p = e.particles(); t = nsteps * dt
    low = scene.particles.x[:, 2] < 0.31                                                            # started below the plane
    assert low.sum() > 50
    assert abs(p["v"][low, 2].mean() - w) < 0.25 * w                                                # carried upwards at the plane's speed
    assert (p["x"][low, 2] - scene.particles.x[low, 2]).mean() > 0.7 * w * t
    assert e.clock()["escaped"] == 0
    e.close()


def test_opt_in_coulomb_friction_reduces_the_tangential_velocity():
    dt = 2e-4
    free = _after_grid_update(_scene(levelset=False), dt)[1]
    e0, g0, _ = _after_grid_update(_scene(), dt)                                                    # reference behaviour (default)
    e1, g1, _ = _after_grid_update(_scene(), dt, coulomb_friction=1)
    z = _node_z(_scene()); inside = (z - 0.31 <= 0.0) & (free["m"] > 0)
    vn = free["v"][:, 2]; hit = inside & (vn < 0)
    vt = free["v"].copy(); vt[:, 2] = 0.0; vtn = np.linalg.norm(vt, axis=1)
    slide = hit & (vtn > -MU * vn)
    assert slide.sum() > 10
    scale = np.abs(free["v"]).max()
    # reference: a sliding node keeps its whole tangential velocity (HS:500-501 does nothing) ...
    assert np.abs(g0["v"][slide] - vt[slide]).max() < 2e-6 * scale
    # ... Coulomb: |v_t| shrinks by mu |v_n|, direction kept
    expected = vt[slide] * ((vtn[slide] + MU * vn[slide]) / vtn[slide])[:, None]
    assert np.abs(g1["v"][slide] - expected).max() < 3e-6 * scale
    stick = hit & (vtn <= -MU * vn)
    assert np.abs(g1["v"][stick]).max() == 0.0 and np.abs(g0["v"][stick]).max() == 0.0
    assert np.abs(g0["vt"] - g1["vt"]).max() < 1e-6 * scale                                         # the pre-friction copy (advection, F update) is the same
    e0.close(); e1.close()


def test_opt_in_mass_floor_of_the_dt_rule():
    """vmax_min_mass_fraction = f: max|v| runs over nodes heavier than f * (one particle's mass); 0 = the reference rule (every m > 0)."""
    dt = 2e-4; scene = _scene(); mp = float(scene.particles.m[0])
    e0, g0, c0 = _after_grid_update(_scene(), dt)
    e1, g1, c1 = _after_grid_update(_scene(), dt, vmax_min_mass_fraction=0.05)
    free = _after_grid_update(_scene(levelset=False), dt)[1]                                        # max|v| is taken BEFORE the collider (HS:878 precedes :899)
    speed = np.linalg.norm(free["v"], axis=1)
    assert c0["vmax"] == pytest.approx(speed[free["m"] > 0].max(), rel=1e-5)
    assert c1["vmax"] == pytest.approx(speed[free["m"] > 0.05 * mp].max(), rel=1e-5)
    assert c1["vmax"] <= c0["vmax"]
    assert np.abs(g0["v"] - g1["v"]).max() < 1e-6 * np.abs(g0["v"]).max()                           # only the dt rule's input changes
    e0.close(); e1.close()
