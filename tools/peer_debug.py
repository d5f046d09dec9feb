#!/usr/bin/env python
"""development: in-process peer slab group vs whole context at a given size / capacity (one GPU); prints relative errors"""
import argparse, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
ap = argparse.ArgumentParser()
ap.add_argument("--res", type=int, default=128); ap.add_argument("--steps", type=int, default=24); ap.add_argument("--nslabs", type=int, default=2)
ap.add_argument("--capf", type=float, default=1.3); ap.add_argument("--sort-every", type=int, default=0); ap.add_argument("--graph", type=int, default=1)
ap.add_argument("--vy", type=float, default=1.5); ap.add_argument("--tag", default="")
a = ap.parse_args()
from peer_parity import scene_for, relerr
from anisotropicelastoplasticity_b200 import capi
from anisotropicelastoplasticity_b200.engine import Engine
from anisotropicelastoplasticity_b200.distributed import PeerSlabGroup, SlabPlan, make_gpu_slab_engine
scene = scene_for(a.res)
if a.vy != 1.5: scene.particles.v[:, 1] *= a.vy / 1.5
cells = np.floor(scene.particles.x[:, 1] * a.res).astype(np.int64)
plan = SlabPlan.balanced(cells, a.res, a.nslabs, axis=1)
rf = 300.0 * a.res / 32.0
engs = []
for r in range(a.nslabs):
    eng, local, idx = make_gpu_slab_engine(scene, plan, r, device=0, capacity_factor=a.capf, dt_rate_floor=rf, sort_every=a.sort_every, use_graph=a.graph)
    capi.check(eng.L.aep_set_particle_id_base(eng.h, int(idx[0])), eng.h); eng.upload_particles(local); engs.append(eng)
grp = PeerSlabGroup(engs, migrate_capacity=max(4096, scene.particles.n // 20))
dt = float(np.float32(1e-4))
grp.init()
for e in engs: e.set_fixed_dt(dt)
whole = Engine(scene, dt_rate_floor=rf, sort_every=a.sort_every); whole.init(); whole.set_fixed_dt(dt)
out = {"tag": a.tag, "res": a.res, "n": scene.particles.n, "capf": a.capf, "rounds_env": os.environ.get("AEP_FORCE_ROUNDS"), "graph": a.graph, "sort_every": a.sort_every, "steps": []}
done = 0
for chunk in (1, 1, 2, 4, 8, a.steps):
    k = min(chunk, a.steps - done)
    if k <= 0: break
    grp.run(k); whole.run(k); done += k
    got = grp.gather_particles(); pw = whole.particles()
    ok_ids = bool((got["ids"] == np.arange(scene.particles.n)).all())
    err = {key: relerr(got[key], pw[key]) for key in ("x", "v", "FE")} if ok_ids else {}
    out["steps"].append({"after": done, "ids_ok": ok_ids, **err, "escaped": [e.clock()["escaped"] for e in engs], "sorts": [e.counters()["sorts"] for e in engs], "mig": [e.migration()["sent"] for e in engs]})
print(json.dumps(out))
