"""The C-ABI shared library loads and exports every symbol include/aep_b200.h declares.  No compute calls (CPU box)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "aep_b200.h")).read()
    return sorted(set(re.findall(r"AEP_API\s+[\w\s\*]+?\b(aep_\w+)\s*\(", txt)))


def test_header_and_binding_agree():
    from anisotropicelastoplasticity_b200 import capi
    assert sorted(capi.SYMBOLS) == header_symbols()


def test_library_exports_all_symbols():
    from anisotropicelastoplasticity_b200 import capi
    if not os.path.exists(capi.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    L = C.CDLL(capi.LIB_PATH)
    for s in header_symbols():
        assert hasattr(L, s), f"libaep_b200.so does not export {s}"


def test_config_struct_layout():
    from anisotropicelastoplasticity_b200 import capi
    L = capi.load(); cfg = capi.Config()
    assert L.aep_default_config(C.byref(cfg)) == 0
    assert cfg.material == 1 and cfg.cfl == pytest.approx(0.3) and cfg.gravity == pytest.approx(9.8)      # main.cpp:27, HS:457
    assert cfg.collider_friction == pytest.approx(0.2) and cfg.snow_hardening == pytest.approx(10.0)         # HS:465, HS:267
    assert list(cfg.sand_h) == [35.0, 9.0, 0.2, 10.0] and cfg.dt_rate_floor == 300.0                         # HS:641-644, HS:860
    assert cfg.frame_dt == pytest.approx(1.0 / 60.0) and cfg.slab_axis == -1 and cfg.sort_every == 0
    assert cfg.sort_bricks == 0 and cfg.scatter_strips == 64 and cfg.sort_cost_threshold == 0.06
    # the opt-in departures from the reference are OFF by default; last field reached: layouts agree
    assert cfg.vmax_min_mass_fraction == 0.0 and cfg.coulomb_friction == 0 and cfg.use_graph == 1 and C.sizeof(capi.Config) == 208


def test_no_cpu_fallback():
    """Without a GPU aep_create must fail loudly (AEP_ERR_CUDA) -- there is no CPU path behind the ABI."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from anisotropicelastoplasticity_b200 import capi
    L = capi.load(); cfg = capi.Config(); L.aep_default_config(C.byref(cfg))
    h = C.c_void_p()
    rc = L.aep_create(C.byref(h), C.byref(cfg))
    assert rc == -2 and not h
    assert b"no CPU fallback" in L.aep_last_error(None)
    from anisotropicelastoplasticity_b200 import scenes as sc
    from anisotropicelastoplasticity_b200.engine import Engine
    with pytest.raises(capi.AepError):
        Engine(sc.small_block(res=16, cells=2))
