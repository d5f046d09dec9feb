#!/usr/bin/env python
"""development: is the peer_parity scene reproducible at all?  whole context vs whole context (different sort periods -> different
summation order), per-step error growth and escaped counters"""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from peer_parity import scene_for, relerr
from anisotropicelastoplasticity_b200.engine import Engine
res = int(sys.argv[1]) if len(sys.argv) > 1 else 128
scene = scene_for(res); rf = 300.0 * res / 32.0; dt = float(np.float32(1e-4))
a = Engine(scene, dt_rate_floor=rf); a.init(); a.set_fixed_dt(dt)
b = Engine(scene, dt_rate_floor=rf, sort_every=1); b.init(); b.set_fixed_dt(dt)
done = 0
for k in (1, 1, 2, 4, 8, 8, 8):
    a.run(k); b.run(k); done += k
    pa, pb = a.particles(), b.particles()
    fin = [bool(np.isfinite(p[key]).all()) for p in (pa, pb) for key in ("x", "v", "FE", "FP")]
    print(json.dumps({"res": res, "after": done, "x": relerr(pa["x"], pb["x"]), "v": relerr(pa["v"], pb["v"]), "FE": relerr(pa["FE"], pb["FE"]), "finite": fin,
                      "escaped": [a.clock()["escaped"], b.clock()["escaped"]], "vmax": [a.clock()["vmax"], b.clock()["vmax"]],
                      "maxv_particles": [float(np.abs(pa["v"]).max()), float(np.abs(pb["v"]).max())], "FE_dev": float(np.abs(pa["FE"] - np.eye(3)).max())}))
