"""ctypes binding of the C ABI in include/aep_b200.h (libaep_b200.so, hand-written sm_100a CUDA).

No fallback of any kind: if the library is missing, `load()` raises with the build command; if there is no
sm_100 GPU, `aep_create` fails and `check()` raises with the library's message.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AEP_B200_LIB") or os.path.join(_HERE, "libaep_b200.so")   # override: kernel-variant experiments only
_LIB = None

NUM_STAGES = 11
STAGES = ("sort", "p2g", "forces", "grid", "g2p", "mesh", "halo", "g2p2g", "force_scatter", "forces_list", "g2p_list")
MIGRATE_FLOATS = 44

dp = C.POINTER(C.c_double)
i64p = C.POINTER(C.c_int64)


class Config(C.Structure):
    """struct aep_config."""
    _fields_ = [("device", C.c_int32), ("material", C.c_int32),
                ("grid_min", C.c_double * 3), ("grid_max", C.c_double * 3), ("res", C.c_int32 * 3), ("_pad0", C.c_int32),
                ("cfl", C.c_double), ("gravity", C.c_double), ("collider_friction", C.c_double), ("snow_hardening", C.c_double),
                ("sand_h", C.c_double * 4), ("dt_rate_floor", C.c_double), ("frame_dt", C.c_double),
                ("particle_capacity", C.c_int64), ("slab_axis", C.c_int32), ("slab_lo", C.c_int32), ("slab_hi", C.c_int32),
                ("sort_every", C.c_int32), ("sort_bricks", C.c_int32), ("scatter_strips", C.c_int32), ("sort_cost_threshold", C.c_double),
                ("vmax_min_mass_fraction", C.c_double), ("coulomb_friction", C.c_int32), ("use_graph", C.c_int32)]


class AepError(RuntimeError):
    pass


# every symbol include/aep_b200.h declares (tests check the library exports all of them)
SYMBOLS = (
    "aep_default_config", "aep_create", "aep_destroy", "aep_last_error", "aep_sync", "aep_upload_particles",
    "aep_upload_mesh", "aep_set_levelset_analytic", "aep_set_levelset_samples", "aep_init", "aep_init_begin", "aep_init_volumes", "aep_init_dt", "aep_substep", "aep_run",
    "aep_run_frames", "aep_p2g", "aep_stage_forces", "aep_stage_grid", "aep_stage_g2p", "aep_set_dt", "aep_set_fixed_dt", "aep_get_clock",
    "aep_num_particles", "aep_download_particles", "aep_download_grid", "aep_download_mesh", "aep_download_positions_f32",
    "aep_stats", "aep_kernel_launches", "aep_stream", "aep_profile", "aep_get_timers", "aep_grid_activity", "aep_halo_info",
    "aep_halo_pack", "aep_halo_add", "aep_vmax_get", "aep_vmax_set", "aep_step_forces", "aep_step_grid", "aep_step_g2p",
    "aep_step_p2g", "aep_migrate_extract", "aep_migrate_insert", "aep_migrate_bind", "aep_migrate_extract_begin", "aep_migrate_extract_end",
    "aep_step_p2g_arrivals", "aep_set_particle_id_base", "aep_download_particles_local", "aep_resume", "aep_set_clock",
    "aep_set_escaped", "aep_host_alloc", "aep_host_free", "aep_set_collider_motion", "aep_init_dt_async", "aep_frame_positions_begin", "aep_frame_positions_wait", "aep_get_counters", "aep_get_migration",
    "aep_comm_export", "aep_comm_connect", "aep_comm_connect_local", "aep_group_init", "aep_group_run",
)
COMM_BLOB_BYTES = 256


def load():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise AepError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       f"(nvcc -gencode arch=compute_100a,code=sm_100a).  There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.aep_last_error.restype = C.c_char_p; L.aep_last_error.argtypes = [vp]
    L.aep_create.argtypes = [C.POINTER(vp), C.POINTER(Config)]
    L.aep_default_config.argtypes = [C.POINTER(Config)]
    L.aep_num_particles.restype = C.c_int64; L.aep_num_particles.argtypes = [vp]
    L.aep_kernel_launches.restype = C.c_int64; L.aep_kernel_launches.argtypes = [vp]
    L.aep_stream.restype = vp; L.aep_stream.argtypes = [vp]
    L.aep_upload_particles.argtypes = [vp, C.c_int64] + [dp] * 10 + [C.c_double] * 4
    L.aep_upload_mesh.argtypes = [vp, C.c_int64, C.c_int64, dp, dp, dp, dp, dp, C.POINTER(C.c_int32), dp, dp, dp, dp, dp, dp, dp] + [C.c_double] * 5
    L.aep_set_levelset_analytic.argtypes = [vp, C.c_int, dp]
    L.aep_set_levelset_samples.argtypes = [vp, C.POINTER(C.c_uint8), dp]
    for name in ("aep_destroy", "aep_sync", "aep_init", "aep_init_begin", "aep_init_volumes", "aep_init_dt", "aep_init_dt_async", "aep_substep", "aep_step_forces", "aep_step_grid", "aep_step_g2p", "aep_step_p2g", "aep_frame_positions_wait"):
        getattr(L, name).argtypes = [vp]
    L.aep_run.argtypes = [vp, C.c_int]
    L.aep_run_frames.argtypes = [vp, C.c_int, C.c_int, i64p]
    L.aep_p2g.argtypes = [vp, C.c_int]
    for name in ("aep_stage_forces", "aep_stage_grid", "aep_stage_g2p", "aep_set_dt", "aep_set_fixed_dt"):
        getattr(L, name).argtypes = [vp, C.c_double]
    L.aep_get_clock.argtypes = [vp, dp, dp, dp, C.POINTER(C.c_int32), i64p, dp, i64p]
    L.aep_resume.argtypes = [vp]
    L.aep_set_clock.argtypes = [vp, C.c_double, C.c_double, C.c_double, C.c_int32, C.c_int64]
    L.aep_download_particles.argtypes = [vp] + [dp] * 9
    L.aep_download_grid.argtypes = [vp] + [dp] * 4
    L.aep_download_mesh.argtypes = [vp] + [dp] * 7
    L.aep_download_positions_f32.argtypes = [vp, C.POINTER(C.c_float)]
    L.aep_stats.argtypes = [vp, dp, dp, dp, dp]
    L.aep_profile.argtypes = [vp, C.c_int]
    L.aep_get_timers.argtypes = [vp, dp, i64p]
    L.aep_grid_activity.argtypes = [vp, i64p, i64p]
    L.aep_halo_info.argtypes = [vp, C.c_int, C.c_int, i64p]
    L.aep_halo_pack.argtypes = [vp, C.c_int, C.c_int, vp]
    L.aep_halo_add.argtypes = [vp, C.c_int, C.c_int, vp]
    L.aep_vmax_get.argtypes = [vp, vp]; L.aep_vmax_set.argtypes = [vp, vp]
    L.aep_migrate_extract.argtypes = [vp, vp, vp, C.c_int64, i64p, i64p]
    L.aep_migrate_insert.argtypes = [vp, vp, C.c_int64, vp, C.c_int64]
    L.aep_migrate_bind.argtypes = [vp, vp, vp, C.c_int64, vp]
    L.aep_migrate_extract_begin.argtypes = [vp]
    L.aep_migrate_extract_end.argtypes = [vp, C.c_int64, C.c_int64]
    L.aep_step_p2g_arrivals.argtypes = [vp, C.c_int64]
    L.aep_set_particle_id_base.argtypes = [vp, C.c_int64]
    L.aep_download_particles_local.argtypes = [vp, i64p] + [dp] * 9
    L.aep_set_escaped.argtypes = [vp, C.c_int64]
    L.aep_host_alloc.restype = vp; L.aep_host_alloc.argtypes = [C.c_int64]; L.aep_host_free.argtypes = [vp]; L.aep_host_free.restype = None
    L.aep_set_collider_motion.argtypes = [vp, dp]
    L.aep_frame_positions_begin.argtypes = [vp, C.POINTER(C.c_float)]
    L.aep_get_counters.argtypes = [vp, i64p, i64p, i64p, i64p]
    L.aep_get_migration.argtypes = [vp, i64p, i64p]
    L.aep_comm_export.argtypes = [vp, vp, C.c_int64]
    L.aep_comm_connect.argtypes = [vp, C.c_int, C.c_int, vp]
    L.aep_comm_connect_local.argtypes = [C.POINTER(vp), C.c_int, C.c_int64]
    L.aep_group_init.argtypes = [C.POINTER(vp), C.c_int]
    L.aep_group_run.argtypes = [C.POINTER(vp), C.c_int, C.c_int]
    _LIB = L
    return L


def check(rc, ctx=None):
    if rc != 0:
        msg = load().aep_last_error(ctx)
        raise AepError(f"libaep_b200 error {rc}: {msg.decode() if msg else '?'}")
