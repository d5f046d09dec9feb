"""Diagnostic (not a test; lives under tests/ because it loads the oracle / the fixtures, which only test code may): the engine against the reference-generated fixtures tests/golden/ref_*.npz with the reference's
time steps replayed; prints the error table tests/test_reference_pin.py's GPU tolerances were chosen from.  Never asserts.
Usage on the GPU box: python tests/diag/gpu_refpin_report.py > gpurun_out/refpin.txt"""
import os, sys, time, traceback
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden, relerr
from anisotropicelastoplasticity_b200.engine import Engine

def mom(m, v): return np.asarray(m)[:, None] * np.asarray(v)

t00 = time.time()
for name in ["sand_block", "cloth_only", "snow_sphere", "cloth_sand", "solve_frame", "sand_walls", "snow_block", "sand_corner"]:
    try:
        d, scene = load_golden("ref_" + name)
        e = Engine(scene); e.init()
        out = {"dt0": abs(e.dt - float(d["dt0"])) / float(d["dt0"])}
        dt_prev = float(d["dt0"])
        for dt in map(float, d["dts"]):
            e.stage_forces(dt_prev); e.stage_grid(dt_prev); e.stage_g2p(dt); e.p2g(False); dt_prev = dt
        g = e.grid(); out["gm"] = relerr(g["m"], d["o_gm"]); out["mom"] = relerr(mom(g["m"], g["v"]), mom(d["o_gm"], d["o_gv"]))
        if scene.particles is not None:
            p = e.particles()
            for k in ("x", "v", "FE", "FP", "B"): out[k] = relerr(p[k], d["o_" + k])
            out["q_abs"] = float(np.abs(p["q"] - d["o_q"]).max())
        if scene.mesh is not None:
            m = e.mesh()
            for k in ("vx", "vv", "ex", "ev", "ed"): out[k] = relerr(m[k], d["o_" + k])
        out["escaped"] = e.clock()["escaped"]
        print(name, len(d["dts"]), {k: f"{v:.1e}" for k, v in out.items()}, f"t={time.time()-t00:.1f}s", flush=True)
        e.close()
    except Exception:
        traceback.print_exc(); sys.stdout.flush()
