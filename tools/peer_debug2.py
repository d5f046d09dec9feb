#!/usr/bin/env python
"""development: locate the first divergence between the in-process peer slab group and the whole context (one GPU)"""
import argparse, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
ap = argparse.ArgumentParser()
ap.add_argument("--res", type=int, default=128); ap.add_argument("--steps", type=int, default=3); ap.add_argument("--nslabs", type=int, default=2)
a = ap.parse_args()
from peer_parity import scene_for, relerr
from anisotropicelastoplasticity_b200 import capi
from anisotropicelastoplasticity_b200.engine import Engine
from anisotropicelastoplasticity_b200.distributed import PeerSlabGroup, SlabPlan, make_gpu_slab_engine
scene = scene_for(a.res); res = a.res
cells = np.floor(scene.particles.x[:, 1] * a.res).astype(np.int64)
plan = SlabPlan.balanced(cells, a.res, a.nslabs, axis=1)
rf = 300.0 * a.res / 32.0
engs = []
for r in range(a.nslabs):
    eng, local, idx = make_gpu_slab_engine(scene, plan, r, device=0, dt_rate_floor=rf)
    capi.check(eng.L.aep_set_particle_id_base(eng.h, int(idx[0])), eng.h); eng.upload_particles(local); engs.append(eng)
grp = PeerSlabGroup(engs, migrate_capacity=max(4096, scene.particles.n // 20))
dt = float(np.float32(1e-4))
grp.init()
for e in engs: e.set_fixed_dt(dt)
whole = Engine(scene, dt_rate_floor=rf); whole.init(); whole.set_fixed_dt(dt)
def grids():
    gw = whole.grid(); out = []
    for r, e in enumerate(engs):
        g = e.grid(); lo, hi = plan.bounds[r], plan.bounds[r + 1]
        j = (np.arange(res ** 3) // res) % res
        valid = (j >= max(0, lo - 1)) & (j < min(res, hi + 2))
        for key in ("m", "v", "f", "vt"):
            d = np.abs(g[key][valid] - gw[key][valid]); d = d.reshape(d.shape[0], -1).max(axis=1)
            scale = np.abs(gw[key][valid]).max() + 1e-30
            bad = np.nonzero(~(d <= 1e-4 * scale))[0]
            if len(bad):
                nodes = np.nonzero(valid)[0][bad]
                out.append({"slab": r, "key": key, "nbad": int(len(bad)), "scale": float(scale), "first": [[int(n % res), int((n // res) % res), int(n // res // res), float(d[b])] for n, b in list(zip(nodes, bad))[:6]]})
    return out
print(json.dumps({"bounds": plan.bounds, "init_grid_mismatch": grids()}))
for s in range(a.steps):
    grp.run(1); whole.run(1)
    gm = grids()
    got = grp.gather_particles(); pw = whole.particles()
    ok = bool((got["ids"] == np.arange(scene.particles.n)).all())
    rep = {"step": s + 1, "ids_ok": ok, "escaped": [e.clock()["escaped"] for e in engs], "grid_mismatch": gm}
    if ok:
        for key in ("x", "v", "FE"):
            d = np.abs(got[key] - pw[key]).reshape(scene.particles.n, -1).max(axis=1)
            bad = np.nonzero(~(d <= 1e-3 * (np.abs(pw[key]).max())))[0]
            rep[key] = {"nbad": int(len(bad)), "first": [[int(b), float(d[b]), (pw["x"][b] * res).round(2).tolist()] for b in bad[:6]]}
    print(json.dumps(rep))
