"""The reference's dt rule at scale: dt = cfl / max(rate_floor, max_i |v_i| / h) takes the maximum over EVERY node
with m > 0 (RegularGrid.cpp:188-200), including nodes a particle barely touches: there f/m grows like 1/(1 - f) towards the cell face
(grad w / w), in the reference's fp64 arithmetic as well as here.  Bulk statistics are insensitive to that (tests/test_gpu_parity.py),
the step count per frame is not.  This test runs one whole frame of the bench workload (C5 dam break, developed state) at 64^3 free
running -- the engine on the GPU, the fp64 oracle (pinned to the reference's own code, tests/test_reference_pin.py) on the host -- and
compares the DISTRIBUTIONS of dt and max|v|, not trajectories: fp32 must not make the rule collapse more often than fp64 does.
Measured in the dev container for the fp64 oracle: 165 substeps for the frame, dt at 0.2 % / 4 % / 87 % of the cap cfl / rate_floor
(10th / 50th / 90th percentile) -- the collapse is the reference algorithm's, in double precision."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _scene():
    import bench as B
    from anisotropicelastoplasticity_b200 import scenes as sc
    s = sc.c5_dam_break(res=64)
    z = s.particles.x[:, 2]
    s.particles.v[:, 0] = B.FLOW_U * z / B.FLOW_H; s.particles.v[:, 1] = B.FLOW_V * np.sin(2.0 * np.pi * z / B.FLOW_H)
    return s, B.rate_floor_for(64)


def test_dt_and_vmax_distributions_over_one_frame_engine_vs_fp64():
    from anisotropicelastoplasticity_b200.engine import Engine
    from oracle.oracle_py import Oracle
    scene, rf = _scene()
    dt_cap = 0.3 / rf
    e = Engine(scene, dt_rate_floor=rf); e.init()
    dts_e, vm_e = [], []
    while e.clock()["frame"] < 1 and len(dts_e) < 5000:
        e.substep(); c = e.clock(); dts_e.append(c["dt"]); vm_e.append(0.3 * scene.grid.h.min() / c["dt"])
    scene2, _ = _scene()
    o = Oracle(scene2, threads=0, rate_floor=rf); o.init()
    dts_o, vm_o = [], []
    while o.frame < 1 and len(dts_o) < 5000:
        o.substep(); dts_o.append(o.dt); vm_o.append(0.3 * scene2.grid.h.min() / o.dt)        # the max|v| the rule saw (exact whenever dt is below the cap)
    dts_e, dts_o, vm_e, vm_o = map(np.asarray, (dts_e, dts_o, vm_e, vm_o))
    q = lambda a: np.percentile(a, [10, 50, 90]).tolist()
    print(f"substeps per frame: engine {len(dts_e)}, fp64 {len(dts_o)}; dt/dt_cap percentiles 10/50/90: engine {q(dts_e / dt_cap)}, fp64 {q(dts_o / dt_cap)}; "
          f"vmax percentiles: engine {q(vm_e)}, fp64 {q(vm_o)}")
    assert e.clock()["frame"] == 1 and e.clock()["escaped"] == 0 and abs(e.clock()["inner_t"]) < 1e-12            # the frame was clipped exactly (HS:880-892)
    assert sum(dts_e[:-1]) <= 1.0 / 60.0 + 1e-9
    # the rule is fed by the same kind of nodes in both arithmetics: same order of magnitude of steps per frame, of the typical dt and
    # of the typical max|v|; the collapsed steps (dt < 10 % of the cap) are as frequent
    assert 1 / 3 < len(dts_e) / len(dts_o) < 3.0
    assert 1 / 3 < np.median(dts_e) / np.median(dts_o) < 3.0
    assert 1 / 4 < np.percentile(vm_e, 90) / np.percentile(vm_o, 90) < 4.0                                           # the worst nodes are equally bad
    assert abs((dts_e < 0.1 * dt_cap).mean() - (dts_o < 0.1 * dt_cap).mean()) < 0.25
    e.close()
