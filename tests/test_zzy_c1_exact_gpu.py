"""BASELINE.json configs[0] at its exact size (104 361 particles, 64^3): the engine against the reference's own code run live on the
box's host (oracle/_ref/libaep_ref.so travels with the repo).  Its own file, sorted after the suites that have already run on a
B200: this test was written after the round-1 GPU budget was spent (tolerances from profiles/r1_v11_engine_vs_reference_fixtures.txt)."""
import numpy as np
import pytest

from conftest import relerr
from test_reference_pin import _c1_exact, _reference_run, engine_replay, live, mom


@live
@pytest.mark.gpu
def test_c1_exact_engine_vs_reference():
    """The engine against the reference's own code, run live on the box's host (oracle/_ref travels with the repo), on
    configs[0] at its full size: two passes of the loop body with the reference's time steps replayed."""
    from anisotropicelastoplasticity_b200.engine import Engine
    scene = _c1_exact(); d = _reference_run(scene, 2)
    e = Engine(scene); e.init()
    # dt0 = cfl h / max|v_i| is set by ONE node (here a nearly massless one next to the perturbed block): fp32 p/m there, 1e-5
    assert e.dt == pytest.approx(float(d["dt0"]), rel=1e-5) and relerr(e.particles()["vol"], d["vol_init"]) < 1e-5
    engine_replay(e, d)
    p = e.particles(); g = e.grid()
    assert relerr(g["m"], d["o_gm"]) < 1e-5 and relerr(mom(g["m"], g["v"]), mom(d["o_gm"], d["o_gv"])) < 2e-5
    for k, tol in (("x", 1e-5), ("v", 2e-5), ("FE", 1e-5), ("FP", 1e-5), ("B", 1e-4)):
        assert relerr(p[k], d["o_" + k]) < tol, k
    assert np.abs(p["q"] - d["o_q"]).max() < 5e-5 and e.clock()["escaped"] == 0
    e.close()
