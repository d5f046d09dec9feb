"""Diagnostic (not a test; lives under tests/ because it loads the oracle / the fixtures, which only test code may): per-stage engine-vs-oracle error table on several scenes; never asserts.
Usage on the GPU box: python tests/diag/gpu_report.py > gpurun_out/report.txt"""
import os, sys, time, traceback
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT)
from anisotropicelastoplasticity_b200 import scenes as sc
from anisotropicelastoplasticity_b200.engine import Engine
from oracle.oracle_py import Oracle

def rel(a, b): return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))
def mom(g): return g["m"][:, None] * g["v"]

def stagewise(name, scene):
    print(f"== {name}: particles {scene.particles.n} grid {scene.grid.res}", flush=True)
    e = Engine(scene); o = Oracle(scene, threads=0)
    e.init(); o.init()
    ge, go = e.grid(), o.grid(); pe, po = e.particles(), o.particles()
    print(f" init : grid m {rel(ge['m'], go['m']):.2e} mom {rel(mom(ge), mom(go)):.2e} vol {rel(pe['vol'], po['vol']):.2e} dt {e.dt:.8e} vs {o.dt:.8e} x_roundtrip {np.abs(pe['x']-scene.particles.x).max():.2e}")
    dt0 = o.dt
    e.stage_forces(dt0); o.stage_forces(dt0)
    ge, go = e.grid(), o.grid()
    print(f" force: f {rel(ge['f'], go['f']):.2e}  |f| {np.linalg.norm(go['f']):.3e} max abs diff {np.abs(ge['f']-go['f']).max():.3e}")
    e.stage_grid(dt0); o.stage_grid_update(dt0); vmax_o = o.cfl_condition() * scene.grid.h.min(); o.stage_collide()
    ge, go = e.grid(), o.grid(); act = go["m"] > 1e-12 * go["m"].max()
    print(f" grid : v {rel(ge['v'][act], go['v'][act]):.2e} vt {rel(ge['vt'][act], go['vt'][act]):.2e} vmax {e.clock()['vmax']:.8e} vs {vmax_o:.8e}")
    dt1 = 0.3 / max(300.0, vmax_o / scene.grid.h.min())
    e.stage_g2p(dt1); o.stage_g2p(dt1)
    pe, po = e.particles(), o.particles()
    print(" g2p  : " + " ".join(f"{k} {rel(pe[k], po[k]):.2e}" for k in ("x", "v", "B", "FE", "FP")) + f" q maxabs {np.abs(pe['q']-po['q']).max():.2e} (q max {po['q'].max():.3f})")
    e.p2g(False); o.rebuild_weights(); o.p2g(False)
    ge, go = e.grid(), o.grid()
    print(f" p2g  : grid m {rel(ge['m'], go['m']):.2e} mom {rel(mom(ge), mom(go)):.2e} escaped {e.clock()['escaped']}")
    e.close()

def multistep(name, scene, n):
    e = Engine(scene); o = Oracle(scene, threads=0); e.init(); o.init()
    t0 = time.time(); e.run(n); e.sync(); t1 = time.time()
    for _ in range(n): o.substep()
    t2 = time.time()
    pe, po = e.particles(), o.particles(); st = e.stats(); c = e.clock()
    com, ke, jp = sc.bulk_stats(po["x"], po["v"], scene.particles.m, po["FP"])
    print(f"== {name} {n} substeps: gpu {t1-t0:.2f}s cpu {t2-t1:.2f}s | " + " ".join(f"{k} {rel(pe[k], po[k]):.2e}" for k in ("x", "v", "FE", "FP")) +
          f" | com {st['com']} vs {com} ke {st['ke']:.6e} vs {ke:.6e} jp {st['jp']:.8f} vs {jp:.8f} | t {c['t']+c['inner_t']:.6f} vs {o.time:.6f} frame {c['frame']} vs {o.frame} escaped {c['escaped']} dt {c['dt']:.6e} vs {o.dt:.6e}", flush=True)
    e.close()

def cloth_report():
    from oracle.make_golden import scene_from_dict
    d = dict(np.load(os.path.join(ROOT, "tests", "golden", "cloth_sand.npz"))); scene = scene_from_dict(d, "cloth_sand")
    e = Engine(scene); o = Oracle(scene); e.init(); o.init()
    ge, go = e.grid(), o.grid()
    print(f"== cloth_sand init: grid m {rel(ge['m'], go['m']):.2e} mom {rel(mom(ge), mom(go)):.2e} dt {e.dt:.8e} vs {o.dt:.8e}")
    dt0 = o.dt; e.stage_forces(dt0); o.stage_forces(dt0); ge, go = e.grid(), o.grid()
    print(f" force: f {rel(ge['f'], go['f']):.2e} max abs {np.abs(ge['f']-go['f']).max():.3e} of {np.abs(go['f']).max():.3e}")
    e.stage_grid(dt0); o.stage_grid_update(dt0); vmax_o = o.cfl_condition() * scene.grid.h.min(); o.stage_collide()
    ge, go = e.grid(), o.grid(); act = go["m"] > 1e-12 * go["m"].max()
    print(f" grid : v {rel(ge['v'][act], go['v'][act]):.2e} vt {rel(ge['vt'][act], go['vt'][act]):.2e}")
    dt1 = 0.3 / max(300.0, vmax_o / scene.grid.h.min()); e.stage_g2p(dt1); o.stage_g2p(dt1)
    me, mo = e.mesh(), o.mesh()
    print(" mesh : " + " ".join(f"{k} {rel(me[k], mo[k]):.2e}" for k in ("vx", "vv", "vB", "ex", "ev", "eB", "ed")))
    e.close()


if __name__ == "__main__":
    try: cloth_report()
    except Exception: traceback.print_exc()
    rng = np.random.default_rng(3)
    cases = {
        "sand_small": sc.small_block(material=sc.SAND, res=16, cells=3, seed=7),
        "snow_small": sc.small_block(material=sc.SNOW, res=16, cells=3, seed=8),
        "sand_corner": sc.small_block(material=sc.SAND, res=12, cells=2, seed=9, lo=(0.0, 0.0, 0.0), levelset=False),
    }
    s = sc.c1_sand_block(res=32); sc.perturb_state(s.particles, rng, strain=1e-2, vel=0.3, affine=1.0); cases["sand_c1_32_perturbed"] = s
    s = sc.c2_snow_sphere(res=32); sc.perturb_state(s.particles, rng, strain=1e-2, vel=0.3, affine=1.0); cases["snow_c2_32_perturbed"] = s
    for k, v in cases.items():
        try: stagewise(k, v)
        except Exception: traceback.print_exc()
    for k, fn, n in (("sand_c1_32", lambda: sc.c1_sand_block(res=32), 1), ("sand_c1_32", lambda: sc.c1_sand_block(res=32), 20), ("sand_c1_32", lambda: sc.c1_sand_block(res=32), 200),
                     ("snow_c2_32", lambda: sc.c2_snow_sphere(res=32), 1), ("snow_c2_32", lambda: sc.c2_snow_sphere(res=32), 200)):
        try: multistep(k, fn(), n)
        except Exception: traceback.print_exc()
