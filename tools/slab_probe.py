#!/usr/bin/env python
"""Single-GPU timing probe of one slab context (no neighbours): does a slab's P2G cost per particle depend on the slab bounds?"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench as B
from anisotropicelastoplasticity_b200 import scenes as sc
from anisotropicelastoplasticity_b200.engine import Engine
res = 512; h = 1.0 / res
import itertools
for (lo, hi), strips in itertools.product(((0, 66), (0, 129), (0, 512)), (1, 64, 1024, 8192)):
    x, _ = B.dam_break_positions(res, seed=5, y_cells=(lo, hi)); n = x.shape[0]
    arrs, keep = B.packed_rest_state(x, sc.SAND_RHO * h ** 3 / 8.0, pinned=False); del x
    slab = None if (lo, hi) == (0, 512) else (1, lo, hi)
    e = Engine(B.make_shell_scene(res), device=0, particle_capacity=int(1.25 * n + 65536), slab=slab, dt_rate_floor=B.rate_floor_for(res), scatter_strips=strips)
    e.upload_packed(n, arrs, sc.SAND_E, sc.SAND_NU, 2.5e-2, 7.5e-3); e.init(); e.run(4); e.sync()
    e.profile(True); e.run(6); e.sync(); tm = e.timers(); e.profile(False)
    st = {k: round(v[0] / max(v[1], 1), 3) for k, v in tm.items()}
    print(json.dumps({"strips": strips, "slab": [lo, hi], "n": n, "stage_ms": st, "ns_per_particle": {k: round(1e6 * st[k] / n, 4) for k in ("p2g", "forces", "g2p")}}), flush=True)
    e.close(); del e, arrs, keep
