"""Seeded random scenes for sweeps (tests/test_reference_pin.py on the CPU, tests/diag/gpu_random_report.py on the GPU)."""
import numpy as np

from anisotropicelastoplasticity_b200 import scenes as sc


def random_particle_scene(seed):
    """Anisotropic grid, 30-150 scattered particles anywhere at least 2 cells inside the domain (also isolated ones with mostly massless
    stencils), random F_E / F_P / B / v / q, either material, a ground or wall-corner collider cutting through the cloud, or none."""
    rng = np.random.default_rng(1000 + seed)
    res = rng.integers(8, 15, size=3); mn = rng.uniform(-1.0, 0.5, size=3); mx = mn + rng.uniform(0.8, 1.6, size=3)
    g = sc.GridSpec(mn, mx, res); h = g.h
    n = int(rng.integers(30, 150))
    centre = mn + (2.5 + rng.random(3) * (res - 5)) * h
    x = np.clip(centre + rng.standard_normal((n, 3)) * h * rng.uniform(0.3, 2.0), mn + 2.01 * h, mx - 2.01 * h)
    material = sc.SAND if seed % 2 else sc.SNOW
    E, nu = (sc.SAND_E, sc.SAND_NU) if material == sc.SAND else (sc.SNOW_E, sc.SNOW_NU)
    F = np.eye(3) + 0.05 * rng.standard_normal((n, 3, 3)); FP = np.eye(3) + 0.03 * rng.standard_normal((n, 3, 3))
    ps = sc.Particles(x=x, v=rng.standard_normal((n, 3)), B=0.5 * rng.standard_normal((n, 3, 3)), FE=F, FP=FP, m=rng.uniform(0.5, 2.0, n) * 1e-3,
                      vol=np.ones(n), q=rng.uniform(0.0, 0.5, n), E=E, nu=nu)
    if seed % 3 == 0:
        ls = sc.LevelSetSpec(sc.LS_GROUND, np.array([centre[2] - 0.3 * h[2], 0, 0, 0, 0, 0, 0, 0.0]))
    elif seed % 3 == 1:
        ls = sc.LevelSetSpec(sc.LS_WALL2GROUND, np.array([centre[0] + 0.4 * h[0], centre[1] + 1.3 * h[1], centre[2] - 1.1 * h[2], 0, 0, 0, 0, 0.0]))
    else:
        ls = sc.LevelSetSpec()
    return sc.Scene(f"random_{seed}", g, material, ps, None, ls)


def random_cloth_scene(seed):
    """A sheet in a random orientation (rotated rest frame), stretched / sheared / crumpled vertex positions, normal directors d3 both
    longer and shorter than 1 (both branches of HS:414 and HS:699-716), random shear stiffness and friction angle (0 included: the cone
    collapses), random pinned vertices, with (odd seeds) and without sand above it, ground collider through the sheet or none."""
    rng = np.random.default_rng(2000 + seed)
    res = rng.integers(10, 15, size=3); g = sc.GridSpec(np.zeros(3), np.ones(3) * rng.uniform(0.9, 1.3), res); h = g.h
    n = int(rng.integers(4, 8)); edge = float(h.min()) * rng.uniform(0.7, 1.2)
    Q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
    du, dv = Q[:, 0] * edge, Q[:, 1] * edge
    centre = g.mx * 0.5
    origin = centre - 0.5 * (n - 1) * (du + dv)
    shear = float(rng.choice([0.0, 20.0, 200.0])); ang = float(rng.choice([0.0, 10.0, 40.0]))
    fixed = tuple(int(v) for v in rng.choice(n * n, size=int(rng.integers(0, 3)), replace=False))
    mesh = sc.make_cloth(n, n, origin, du, dv, shear=shear, friction_angle_deg=ang, stiff=float(rng.choice([4e4, 1e3])), fixed_ids=fixed)
    mesh.vx = mesh.vx + 0.08 * edge * rng.standard_normal(mesh.vx.shape)                       # stretched / crumpled
    mesh.vx = np.clip(mesh.vx, 2.01 * h, g.mx - 2.01 * h)
    mesh.vv = 0.5 * rng.standard_normal(mesh.vv.shape); mesh.ev = 0.5 * rng.standard_normal(mesh.ev.shape)
    mesh.vB = 0.3 * rng.standard_normal(mesh.vB.shape); mesh.eB = 0.3 * rng.standard_normal(mesh.eB.shape)
    mesh.ed[0] = mesh.vx[mesh.faces[:, 1]] - mesh.vx[mesh.faces[:, 0]]; mesh.ed[1] = mesh.vx[mesh.faces[:, 2]] - mesh.vx[mesh.faces[:, 0]]
    mesh.ed[2] = mesh.ed[2] * rng.uniform(0.8, 1.2, size=(mesh.nf, 1)) + 0.1 * rng.standard_normal((mesh.nf, 3))
    ps = None
    if seed % 2:
        m = int(rng.integers(20, 80)); x = np.clip(centre + np.array([0, 0, 1.5 * h[2]]) + rng.standard_normal((m, 3)) * h, 2.01 * h, g.mx - 2.01 * h)
        ps = sc.Particles(x=x, v=rng.standard_normal((m, 3)), B=0.3 * rng.standard_normal((m, 3, 3)), FE=np.eye(3) + 0.03 * rng.standard_normal((m, 3, 3)),
                          FP=np.tile(np.eye(3), (m, 1, 1)), m=np.full(m, 2e-3), vol=np.ones(m), q=np.zeros(m), E=sc.SAND_E, nu=sc.SAND_NU)
    ls = sc.LevelSetSpec(sc.LS_GROUND, np.array([centre[2] - 0.2 * h[2], 0, 0, 0, 0, 0, 0, 0.0])) if seed % 3 else sc.LevelSetSpec()
    return sc.Scene(f"random_cloth_{seed}", g, sc.SAND, ps, mesh, ls)


def degenerate_scene(material):
    """Particles whose F_E is special: identity, two equal singular values, a pure rotation, a reflection (det < 0), strong compression,
    strong anisotropy, nearly singular, uniform dilation; particles exactly on a grid node, on a cell face and on a cell edge."""
    scene = sc.small_block(material=material, res=14, cells=3, seed=55, lo=(0.36, 0.36, 0.3))
    p = scene.particles; rng = np.random.default_rng(56); h = 1.0 / 14
    Q, _ = np.linalg.qr(rng.standard_normal((3, 3))); Q2, _ = np.linalg.qr(rng.standard_normal((3, 3)))
    special = [np.eye(3), Q @ np.diag([1.05, 1.05, 0.9]) @ Q2.T, Q, Q @ np.diag([1.0, 1.0, -1.0]) @ Q.T, 0.6 * np.eye(3) + 0.01 * rng.standard_normal((3, 3)),
               Q @ np.diag([1.6, 1.0, 0.7]) @ Q2.T, Q @ np.diag([1.0, 0.9, 1e-3]) @ Q2.T, np.diag([1.02, 1.02, 1.02])]
    for k, F in enumerate(special):
        p.FE[k] = F
    p.x[10] = np.array([6, 6, 6]) * h                      # exactly on a node
    p.x[11] = np.array([6.0, 6.37, 5.81]) * h              # on a cell face
    p.x[12] = np.array([6.5, 7.0, 6.0]) * h                # on a cell edge
    return scene, len(special)
