"""Diagnostic (not a test; lives under tests/ because it loads the reference library, which only test code may): the engine against
the reference's own code (oracle/_ref, run live on the box's host) over the seeded random scenes of tests/random_scenes.py -- init + 2
substeps, the reference's time steps replayed through aep_stage_*.  Prints the error table per scene and the worst case per
quantity; never asserts.  Usage on the GPU box: python tests/diag/gpu_random_report.py > gpurun_out/random.txt"""
import os, sys, traceback
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import relerr
from random_scenes import degenerate_scene, random_cloth_scene, random_particle_scene
from anisotropicelastoplasticity_b200 import scenes as sc
from anisotropicelastoplasticity_b200.engine import Engine
from oracle.ref_py import Reference

def mom(g): return g["m"][:, None] * g["v"]

worst = {}
def note(tag, out):
    print(tag, {k: f"{v:.1e}" for k, v in out.items()}, flush=True)
    for k, v in out.items():
        if v > worst.get(k, (0.0, ""))[0]: worst[k] = (v, tag)

for kind, gen, count in (("particles", random_particle_scene, 40), ("cloth", random_cloth_scene, 24),
                         ("degenerate", lambda i: degenerate_scene(sc.SAND if i else sc.SNOW)[0], 2)):
    for seed in range(count):
        try:
            scene = gen(seed)
            r = Reference(scene); r.init(); dt0 = float(r.dt)
            e = Engine(scene); e.init()
            out = {"dt0": abs(e.dt - dt0) / dt0}
            dts = [r.substep() for _ in range(2)]
            dt_lag = dt0
            for dt in dts:
                e.stage_forces(dt_lag); e.stage_grid(dt_lag); e.stage_g2p(float(dt)); e.p2g(False); dt_lag = float(dt)
            ge, gr = e.grid(), r.grid()
            out["gm"] = relerr(ge["m"], gr["m"]); out["mom"] = relerr(mom(ge), mom(gr))
            if scene.particles is not None:
                pe, pr = e.particles(), r.particles()
                for k in ("x", "v", "FE", "FP", "B"): out[k] = relerr(pe[k], pr[k])
                out["q_abs"] = float(np.abs(pe["q"] - pr["q"]).max())
            if scene.mesh is not None:
                me, mr = e.mesh(), r.mesh()
                for k in ("vx", "vv", "ex", "ev", "ed"): out["m_" + k] = relerr(me[k], mr[k])
            out["escaped"] = float(e.clock()["escaped"])
            note(f"{kind}[{seed}]", out)
            e.close()
        except Exception:
            traceback.print_exc(); sys.stdout.flush()
print("WORST", {k: (f"{v:.1e}", t) for k, (v, t) in worst.items()})
