#!/usr/bin/env python
"""bench.py -- particle-substeps/s of the full MPM substep (HybridSolver.cpp:867-1032) on N B200s.

  python bench.py --gpus N --steps K --warmup W            # this engine (libaep_b200.so through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path: its own sources compiled against a
                                                           # MiniEigen stand-in (oracle/_ref), else the oracle port

Workload (config.workload): BASELINE.json configs[4], the synthetic 64M-particle sand dam break on a 512^3 grid --
the configuration the headline target is quoted on; it fits one B200 (~33 GB), so it is also the N=1 workload and
the N>1 runs are STRONG scaling of the same scene split into y-slabs.  `--res R` shrinks it (particles ~ R^3).

A "step" is one substep over all particles.  `value` = particles * K / (device time of K substeps, max over ranks),
state resident in HBM.  `e2e` = the same through the host-facing call sequence of HybridSolver::solve: upload of the
fp64 host state (reference layouts, pinned), K substeps, float32 position download (the per-frame OBJ payload).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from anisotropicelastoplasticity_b200 import scenes as sc  # noqa: E402

METRIC = "particle_substeps_per_sec"
UNIT = "particle-substeps/s"
# algorithmic bytes, SURVEY.md 8(d): per particle-substep 340 B (sand) / 376 B (snow); per active grid node 172 B
BYTES_PARTICLE = {sc.SAND: 340.0, sc.SNOW: 376.0}
BYTES_NODE = 172.0
# per-stage split of the same model (sand; snow adds 36 B to forces)
# the fused kernel of a substep does the G2P of this substep and the P2G of the next: its algorithmic bytes are the sum of the two rows
# "forces" = the gather / stress kernel (its half of the stage's node bytes); the scatter half of that stage (k_force_scatter) only has the 12 B
# per node of its reductions as algorithmic bytes -- the A matrices it reads are a device-internal round trip, not credited
STAGE_BYTES = {"forces": (52.0, 12.0), "force_scatter": (0.0, 12.0), "g2p": (224.0, 24.0), "p2g": (64.0, 44.0), "g2p2g": (288.0, 68.0), "grid": (0.0, 52.0), "sort": (0.0, 0.0)}
KERNEL_OF = {"forces": "k_forces<SPLIT> (grad v gather + stress)", "force_scatter": "k_force_scatter", "g2p2g": "k_g2p2g (fused, AEP_FUSED=1)", "g2p": "k_g2p2g<SCATTER=0> (G2P)",
             "p2g": "k_p2g", "grid": "k_grid_update", "sort": "cub radix sort + k_reorder"}
TRAFFIC_OF = {"forces": "k_forces", "force_scatter": "k_force_scatter", "g2p2g": "k_g2p2g", "g2p": "k_g2p2g", "p2g": "k_p2g"}


def kernel_roofline(stage_ms, n, nodes, peak, peak_kind):
    """The dominant KERNEL (per-kernel CUDA-event times of the profiled pass; the list passes over strays are separate launches with
    their own timers) against the measured HBM copy bandwidth, plus the force stage as a whole (gather + list pass + scatter:
    the SURVEY 8(d) unit of 52 B per particle + 24 B per node)."""
    dom = max(("forces", "g2p2g", "g2p", "p2g", "grid"), key=lambda k: stage_ms.get(k, 0.0))
    bp, bn = STAGE_BYTES[dom]
    dom_bytes = bp * n + bn * nodes
    achieved = dom_bytes / (stage_ms[dom] * 1e-3) / 1e9 if stage_ms[dom] > 0 else 0.0
    f_ms = stage_ms.get("forces", 0.0) + stage_ms.get("forces_list", 0.0) + stage_ms.get("force_scatter", 0.0)
    f_bytes = 52.0 * n + 24.0 * nodes
    return dom, {"bound": "hbm", "kernel": KERNEL_OF[dom], "achieved": achieved, "peak": peak, "peak_kind": peak_kind + " copy bandwidth, MEASURED_PEAKS.json", "unit": "GB/s",
                 "frac": achieved / peak, "traffic": ncu_traffic(TRAFFIC_OF.get(dom, ""), n), "algorithmic_bytes_per_launch": dom_bytes, "launch_ms": stage_ms[dom],
                 "stage_ms": stage_ms,
                 "force_stage": {"kernels": "k_forces<SPLIT> + its list pass + k_force_scatter", "algorithmic_bytes": f_bytes, "ms": f_ms,
                                 "achieved_gbs": f_bytes / (f_ms * 1e-3) / 1e9 if f_ms > 0 else 0.0, "frac": (f_bytes / (f_ms * 1e-3) / 1e9 / peak) if f_ms > 0 else None}}


def ncu_traffic(kernel, particles):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel`, from the committed `ncu --set full` capture of this
    build (profiles/ncu_traffic.json, written from the capture by tools/ncu_traffic.py).  The capture is of the bench workload at
    512^3; other sizes get it scaled per particle.  None when the file does not cover the kernel."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f)
        k = t["kernels"][kernel]
        return (k["dram_bytes_read"] + k["dram_bytes_write"]) * particles / t["particles"]
    except Exception:
        return None


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md): NVML from a thread every 5 ms (the timed region
    of the 8-GPU run is ~50 ms, nvidia-smi's loop cannot go below 100 ms), nvidia-smi as the fall-back."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        import threading
        self.p = None; self.f = None; self.th = None; self.samples = []; self.reasons = set(); self.max_mhz = None; self._stop = threading.Event()
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[device]) if vis and all(t.strip().isdigit() for t in vis.split(",")) else device
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8), "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20), "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")

            def loop():
                while not self._stop.is_set():
                    try:
                        self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                        r = int(get_reasons(h))
                        for k, bit in names.items():
                            if r & bit:
                                self.reasons.add(k)
                    except Exception:
                        pass
                    self._stop.wait(0.005)
            self.th = threading.Thread(target=loop, daemon=True); self.th.start()
            return
        except Exception:
            self.th = None
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(device)],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.th is not None:
            self._stop.set(); self.th.join(timeout=2)
            if self.samples:
                busy = [s for s in self.samples if s >= 0.5 * max(self.samples)]
                out = {"sm_mhz": statistics.median(busy), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples), "source": "nvml, 5 ms period"}
            return out
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush(); self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            busy = [s for s in sm if s >= 0.5 * max(sm)]
            out = {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi, 100 ms period"}
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


def dam_break_rows(res):
    """Lattice of the C5 column (scenes.c5_dam_break: 8 particles per cell, x, z in [2h, 0.25), y in [2h, 1-2h)) as sub-cell index
    ranges per axis: (i0, i1), (j0, j1), (k0, k1) in units of h/2."""
    h = 1.0 / res
    c0 = np.array([2, 2, 2]); c1 = np.array([int(np.ceil(0.25 / h - 1e-9)), res - 2, int(np.ceil(0.25 / h - 1e-9))])
    return [(2 * int(a), 2 * int(b)) for a, b in zip(c0, c1)]


def dam_break_positions(res, seed=5, y_cells=None):
    """C5 positions only (memory-lean).  Every y sub-row of the lattice draws its jitter from its own stream (seed, row), so a rank
    of the N-GPU run generates exactly its rows of the SAME scene the one-GPU run holds: `y_cells` = (c0, c1) restricts to cells
    [c0, c1) along y.  Returns (positions, index of the first particle in the whole scene's order = global id base)."""
    h = 1.0 / res; sub = 0.5 * h
    (i0, i1), (j0, j1), (k0, k1) = dam_break_rows(res)
    per_row = (i1 - i0) * (k1 - k0)
    ja, jb = (j0, j1) if y_cells is None else (max(j0, 2 * y_cells[0]), min(j1, 2 * y_cells[1]))
    out = np.empty((max(0, jb - ja) * per_row, 3))
    k, i = np.meshgrid(np.arange(k0, k1), np.arange(i0, i1), indexing="ij")
    ik = np.stack([i.ravel(), np.zeros(per_row), k.ravel()], axis=1).astype(np.float64)
    for n, j in enumerate(range(ja, jb)):
        ik[:, 1] = j
        jit = np.random.default_rng([seed, j]).random((per_row, 3))
        out[n * per_row:(n + 1) * per_row] = (ik + 0.05 + 0.9 * jit) * sub
    return out, (ja - j0) * per_row


def packed_rest_state(x, mass, pinned=False):
    """Host arrays in the reference's layouts for particles at rest (F = I, v = B = 0), built without (N,3,3) temporaries."""
    import torch
    n = x.shape[0]

    def buf(shape, fill=0.0):
        t = torch.empty(shape, dtype=torch.float64, pin_memory=pinned)
        a = t.numpy(); a[...] = fill
        return t, a
    keep = []
    xs_t, xs = buf((3, n)); xs[...] = x.T
    v_t, v = buf((3, n)); b1_t, b1 = buf((3, n)); b2_t, b2 = buf((3, n)); b3_t, b3 = buf((3, n))
    fe_t, fe = buf((n, 9)); fe[:, 0] = 1.0; fe[:, 4] = 1.0; fe[:, 8] = 1.0
    fp_t, fp = buf((n, 9)); fp[...] = fe
    m_t, m = buf((n,), mass); vol_t, vol = buf((n,), 1.0); q_t, q = buf((n,))
    keep = [xs_t, v_t, b1_t, b2_t, b3_t, fe_t, fp_t, m_t, vol_t, q_t]
    return [xs, v, b1, b2, b3, fe, fp, m, vol, q], keep


FLOW_U, FLOW_V, FLOW_H = 2.0, 1.0, 0.25


def flowing_packed(arrs, chunk=1 << 22):
    """The workload's DEVELOPED state (closed form): the column sheared the way a collapsing dam-break column is -- v_x = U z / H (the
    top runs ahead), v_y = V sin(2 pi z / H) (particles cross the y-slab boundaries of the N-GPU runs in both directions), U = 2 m/s,
    V = 1 m/s, H = 0.25 m.  With dt = 0.3 / rate_floor a particle moves up to 0.064 cells per substep: within the warm-up every
    particle is straining at 8 /s, cohesionless sand at zero confining pressure yields on every substep (the Drucker-Prager projection
    and the F_P update run for every particle), particles change cells continuously, the adaptive re-sort fires every few substeps
    and slab contexts exchange particles -- what a running simulation pays, inside the timed region."""
    xs, v = arrs[0], arrs[1]
    n = xs.shape[1]
    for p0 in range(0, n, chunk):
        p1 = min(n, p0 + chunk)
        z = xs[2, p0:p1]
        v[0, p0:p1] = FLOW_U * z / FLOW_H
        v[1, p0:p1] = FLOW_V * np.sin(2.0 * np.pi * z / FLOW_H)


def perturb_packed(arrs, strain, seed=17, chunk=1 << 22):
    """Development option (--perturb): a state in which every branch of the substep works -- random elastic strains of scale
    `strain` (the SVDs need 2-3 sweeps, sand yields), velocities of scale 30*strain m/s (particles change cells, the re-sort
    policy fires).  The headline line is measured on the rest state the workload names; this shows what a flowing state costs."""
    rng = np.random.default_rng(seed)
    xs, v, b1, b2, b3, fe, fp, m, vol, q = arrs
    n = fe.shape[0]
    for p0 in range(0, n, chunk):
        p1 = min(n, p0 + chunk)
        fe[p0:p1] += strain * rng.standard_normal((p1 - p0, 9))
        v[:, p0:p1] += 30.0 * strain * rng.standard_normal((3, p1 - p0))


def rate_floor_for(res):
    """dt = cfl / max(rate_floor, vmax/h) (HybridSolver.cpp:860,878).  The reference's literal 300 (dt <= 1e-3 s) is tuned to its
    own coarse grid (h = 0.044, main.cpp:53-69): with sand's p-wave speed ~19 m/s it gives c dt / h = 0.43 there, but 9.8 on a
    512^3 unit box, where the explicit update blows up.  The floor is a config value of the ABI (aep_config.dt_rate_floor);
    the bench scales it with resolution so that c dt / h stays at 0.61 (the value of the 32^3 parity scenes)."""
    return 300.0 * res / 32.0


def make_shell_scene(res):
    """Grid + level set of C5 without particles."""
    g = sc.GridSpec(np.zeros(3), np.ones(3), np.array([res] * 3))
    h = g.h; e = 1e-4 * h[0]
    ls = sc.LevelSetSpec(sc.LS_BOX, np.array([2 * h[0] - e, 2 * h[1] - e, 2 * h[2] - e, 1 - 2 * h[0] + e, 1 - 2 * h[1] + e, 1 - 2 * h[2] + e, 0, 0.0]))
    return sc.Scene("C5_dam_break", g, sc.SAND, None, None, ls)


REF_NOTE = ("the reference's own unmodified sources (HybridSolver.cpp etc.) compiled against the MiniEigen stand-in "
            "(oracle/_ref/libaep_ref.so; this image has no Eigen), serial like the reference")


def _timed_substeps(sim, seconds_budget, max_steps, warmup):
    for _ in range(warmup):
        sim.substep()
    n = 0; t0 = time.perf_counter()
    while True:
        sim.substep(); n += 1
        el = time.perf_counter() - t0
        if el > seconds_budget or n >= max_steps:
            return n, el


def _stage_share(sim):
    """Where the CPU run spends its time, by stage of the loop body (weights = evaluateInterpolationWeights_, HS:18-97: the
    reference's own triplet loops and the four setFromTriplets; forces HS:252-458; grid HS:725-737,460-551; g2p HS:739-825,940-951,
    553-723; p2g HS:113-250)."""
    t = sim.timers(); tot = sum(t.values()) or 1.0
    return {k: round(float(v) / tot, 3) for k, v in t.items()}


def cpu_baseline(threads, seconds_budget=20.0, res=64):
    """The reference's CPU path on a bounded sample of the same workload: the C5 dam break at res^3 (same particles per
    cell, same material, same collider).  kind "reference" = the reference's own code (oracle/_ref, built in the dev container
    from /root/reference and shipped as a .so); the oracle port's figure on the same sample rides along under "port" (it is
    ~7x faster than the reference because it never builds the sparse weight matrices).  Falls back to the port alone only
    if the reference library is missing."""
    from oracle import ref_py
    from oracle.oracle_py import Oracle
    scene = sc.c5_dam_break(res=res)
    o = Oracle(scene, threads=threads, rate_floor=rate_floor_for(res)); o.init()
    used = o.L.orc_get_threads(o.h)
    have_ref = ref_py.available()
    n, el = _timed_substeps(o, seconds_budget / 2 if have_ref else seconds_budget, 200, 2)
    port = {"value": scene.particles.n * n / el, "unit": UNIT, "cores": int(used), "kind": "port",
            "sample": f"C5 dam break at {res}^3 grid, {scene.particles.n} particles, {n} substeps in {el:.1f} s (fp64 oracle, oracle/mpm_oracle.cpp)"}
    if not have_ref:
        return port
    r = ref_py.Reference(scene, rate_floor=rate_floor_for(res)); r.init()
    n, el = _timed_substeps(r, seconds_budget, 50, 1)
    return {"value": scene.particles.n * n / el, "unit": UNIT, "cores": 1, "kind": "reference",
            "sample": f"C5 dam break at {res}^3 grid, {scene.particles.n} particles, {n} substeps in {el:.1f} s; {REF_NOTE}",
            "stage_share": _stage_share(r), "port": port}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores -- its own code (oracle/_ref) when
    that library is present, else the oracle port with all host threads.  The reference is serial, so one core is
    all the host threads it can use."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref_py
    from oracle.oracle_py import Oracle
    res = args.ref_res
    scene = sc.c5_dam_break(res=res)
    if ref_py.available():
        o = ref_py.Reference(scene, rate_floor=rate_floor_for(res)); kind = "reference"; used = 1; what = REF_NOTE
    else:
        o = Oracle(scene, threads=0, rate_floor=rate_floor_for(res)); kind = "port"; used = o.L.orc_get_threads(o.h); what = "fp64 oracle port, all host threads"
    o.init()
    for _ in range(args.warmup):
        o.substep()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.substep()
    el = time.perf_counter() - t0
    val = scene.particles.n * args.steps / el
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(args, res_override=res, note="bounded sample of the workload: same scene at a smaller grid"),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": int(used), "kind": kind,
                             "sample": f"C5 dam break at {res}^3 grid, {scene.particles.n} particles, {args.steps} substeps; {what}",
                             "stage_share": _stage_share(o)},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(args, res_override=None, note=None, n_particles=None):
    res = res_override or args.res
    what = ("developed state: column shearing at v_x = 2 z/H m/s, v_y = sin(2 pi z/H) m/s -- every particle yields, changes cells, re-sorts and slab migration inside the timed region"
            if getattr(args, "state", "flowing") == "flowing" else "rest state (F = I, v = 0)")
    cfg = {"workload": f"C5 synthetic sand dam break (Drucker-Prager), 8 particles/cell, {res}^3 grid; {what}", "grid": [res] * 3, "state": getattr(args, "state", "flowing"),
           "material": "sand", "collider": "box level set", "sort_every": args.sort_every, "timestep": f"reference rule dt = 0.3 / max(rate_floor, vmax/h) on device, rate_floor = {rate_floor_for(res):g} (300 scaled by res/32 for stability)",
           "l2": "inputs (particle state >> 126 MB L2) larger than L2; no explicit flush"}
    if n_particles is not None:
        cfg["particles"] = int(n_particles)
    if note:
        cfg["note"] = note
    return cfg


def transfers_roofline(stage_ms, n, nodes, peak):
    """The two transfers BASELINE.json's north star singles out ("P2G+G2P at 50% or more of the HBM roofline per GPU"): their
    algorithmic bytes (SURVEY.md 8d split) over the sum of their stage times."""
    ms = stage_ms.get("p2g", 0.0) + stage_ms.get("g2p", 0.0) + stage_ms.get("g2p_list", 0.0) + stage_ms.get("g2p2g", 0.0)
    b = sum(STAGE_BYTES[k][0] * n + STAGE_BYTES[k][1] * nodes for k in ("p2g", "g2p"))
    gbs = b / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
    return {"algorithmic_bytes": b, "ms": ms, "achieved_gbs": gbs, "frac": gbs / peak if peak else None}


def set_state(arrs, args):
    """rest (F = I, v = 0: what packed_rest_state built) or the developed state of flowing_packed"""
    arrs[1][...] = 0.0
    if args.state == "flowing":
        flowing_packed(arrs)
    if args.perturb > 0.0:
        perturb_packed(arrs, args.perturb)


def timed_run(eng, stream, args, n, steps=None, warmup=None):
    """W warm-up substeps, then K substeps between two CUDA events on the engine's stream (re-sorts included: they are part of what a
    running simulation pays).  Returns ms and what happened inside the timed region."""
    import torch
    steps = args.steps if steps is None else steps; warmup = args.warmup if warmup is None else warmup
    eng.run(warmup); eng.sync()
    c0 = eng.counters(); k0 = eng.clock(); launches0 = eng.kernel_launches
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); ev0.record(stream)
    eng.run(steps)
    ev1.record(stream); torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    c1 = eng.counters(); k1 = eng.clock()
    sim_s = (k1["t"] + k1["inner_t"]) - (k0["t"] + k0["inner_t"])
    return {"ms": ms, "steps": steps, "launches": eng.kernel_launches - launches0, "sorts": c1["sorts"] - c0["sorts"],
            "moved_fraction_since_last_sort": c1["moved_since_sort"] / max(1, n), "clock": k1, "simulated_seconds": sim_s,
            "simulated_seconds_per_wall_second": sim_s / (ms * 1e-3) if ms > 0 else 0.0}


def run_engine(args):
    import torch
    from anisotropicelastoplasticity_b200.engine import Engine
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 or args.gpus > 1:
        from anisotropicelastoplasticity_b200 import distributed as dist_mod
        return dist_mod.bench_main(args, workload_config, ClockSampler, measured_peak_gbs, METRIC, UNIT)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this engine has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    res = args.res
    t_gen = time.perf_counter()
    x, _ = dam_break_positions(res)
    n = x.shape[0]
    mass = sc.SAND_RHO * (1.0 / res) ** 3 / 8.0
    arrs, keep = packed_rest_state(x, mass, pinned=True)
    del x
    set_state(arrs, args)
    t_gen = time.perf_counter() - t_gen
    shell = make_shell_scene(res)
    rate_floor = rate_floor_for(res)
    mat = (sc.SAND_E, sc.SAND_NU, 2.5e-2, 7.5e-3)
    # one context, created and sized once (23 GB of cudaMalloc is setup, not a step); every leg below uploads into it
    eng = Engine(shell, device=local, particle_capacity=n, dt_rate_floor=rate_floor, sort_every=args.sort_every, sort_bricks=args.sort_bricks,
                 sort_cost_threshold=args.sort_threshold)
    stream = torch.cuda.ExternalStream(eng.stream, device=local)
    eng.upload_packed(n, arrs, *mat); eng.init()
    if args.pin_dt > 0.0:
        eng.set_fixed_dt(args.pin_dt)
    # ---- timed region: K substeps of the developed state, device events on the engine's stream
    sampler = ClockSampler(local)
    tr = timed_run(eng, stream, args, n)
    clocks = sampler.stop()
    ms, launches, clk = tr["ms"], tr["launches"], tr["clock"]
    value = n * args.steps / (ms * 1e-3)
    # ---- per-stage device time (separate pass with 2 events per stage) -> dominant kernel roofline
    blocks, nodes = eng.grid_activity()
    eng.profile(True); eng.run(max(3, min(args.steps, 10))); eng.sync(); tm = eng.timers(); eng.profile(False)
    peak, peak_kind = measured_peak_gbs()
    stage_ms = {k: (v[0] / v[1] if v[1] else 0.0) for k, v in tm.items()}
    # dominant kernel = the launch that costs most PER SUBSTEP; the re-sort (no algorithmic bytes) is overhead inside `value`,
    # reported as sorts_in_timed_region / amortised_sort_ms, not a candidate
    dom, roofline = kernel_roofline(stage_ms, n, nodes, peak, peak_kind)
    sub_bytes = BYTES_PARTICLE[sc.SAND] * n + BYTES_NODE * nodes
    sub_gbs = sub_bytes / (ms * 1e-3 / args.steps) / 1e9
    roofline["substep"] = {"algorithmic_bytes": sub_bytes, "achieved_gbs": sub_gbs, "frac": sub_gbs / peak, "active_nodes": nodes, "active_blocks": blocks}
    roofline["p2g_g2p"] = transfers_roofline(stage_ms, n, nodes, peak)
    state = {"state": args.state, "sorts_in_timed_region": int(tr["sorts"]), "amortised_sort_ms": tr["sorts"] * stage_ms.get("sort", 0.0) / args.steps,
             "moved_fraction_since_last_sort": tr["moved_fraction_since_last_sort"], "simulated_seconds_per_wall_second": tr["simulated_seconds_per_wall_second"]}
    if args.quick:
        print(json.dumps({"metric": METRIC, "value": value, "ms_per_step": ms / args.steps, "sort_every": args.sort_every, "stage_ms": stage_ms,
                          "gpu_launches": int(launches), "sim": clk, "active_nodes": nodes, "particles": n, "perturb": args.perturb, "running": state,
                          "p2g_g2p_frac": roofline["p2g_g2p"]["frac"], "dom": dom, "dom_frac": roofline["frac"]}))
        return
    # ---- e2e: host fp64 state -> device, K substeps, f32 positions back (HybridSolver::solve's host-visible traffic)
    out_t = torch.empty((n, 3), dtype=torch.float32, pin_memory=True)
    import ctypes as C
    from anisotropicelastoplasticity_b200 import capi
    eng.sync(); torch.cuda.synchronize(); t0 = time.perf_counter()
    eng.upload_packed(n, arrs, *mat)
    eng.init()
    eng.run(args.steps)
    capi.check(eng.L.aep_download_positions_f32(eng.h, C.cast(out_t.data_ptr(), C.POINTER(C.c_float))), eng.h)
    t_e2e = time.perf_counter() - t0
    h2d = 36 * 8 * n; d2h = 12 * n
    e2e = {"value": n * args.steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d / args.steps, "d2h_bytes_per_step": d2h / args.steps,
           "seconds": t_e2e, "what": "on a context created and sized beforehand: aep_upload_particles(fp64 host, pinned) + aep_init + K substeps + aep_download_positions_f32"}
    assert np.isfinite(out_t.numpy()).all()
    # ---- secondary: the same scene at rest (round 1's headline state: one SVD sweep, nobody yields, nobody changes cell)
    secondary = None
    if args.state != "rest":
        arrs[1][...] = 0.0
        eng.upload_packed(n, arrs, *mat); eng.init()
        tr2 = timed_run(eng, stream, args, n)
        secondary = {"workload": "same scene at rest (F = I, v = 0)", "value": n * args.steps / (tr2["ms"] * 1e-3), "ms_per_step": tr2["ms"] / args.steps,
                     "sorts_in_timed_region": int(tr2["sorts"]), "simulated_seconds_per_wall_second": tr2["simulated_seconds_per_wall_second"]}
    eng.close()
    cpu = cpu_baseline(threads=1) if not args.no_cpu else None
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, n_particles=n), "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks, "running": state, "secondary": secondary,
            "sim": {"dt": clk["dt"], "t": clk["t"] + clk["inner_t"], "escaped": clk["escaped"], "vmax": clk["vmax"]},
            "setup_s": {"generate": t_gen}}
    print(json.dumps(line))


def run_engine_config(args):
    """--config C1..C4 (development; the contract line is the default C5 run): one of BASELINE.json's other configurations through the
    same engine, state resident in HBM, K substeps timed with CUDA events on the engine's stream.  The unit count of a cloth scene is
    Np + Nv + Nf (vertices and element centroids are transferred like particles, HS:838-849)."""
    import torch
    from anisotropicelastoplasticity_b200.engine import Engine
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this engine has no CPU path")
    local = int(os.environ.get("LOCAL_RANK", "0")); torch.cuda.set_device(local)
    t_gen = time.perf_counter(); scene = sc.CONFIGS[args.config](); t_gen = time.perf_counter() - t_gen
    units = (scene.particles.n if scene.particles is not None else 0) + (scene.mesh.nv + scene.mesh.nf if scene.mesh is not None else 0)
    res = int(max(scene.grid.res))
    eng = Engine(scene, device=local, dt_rate_floor=rate_floor_for(res), sort_every=args.sort_every); eng.init()
    stream = torch.cuda.ExternalStream(eng.stream, device=local)
    eng.run(args.warmup); eng.sync()
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    launches0 = eng.kernel_launches
    torch.cuda.synchronize(); ev0.record(stream); eng.run(args.steps); ev1.record(stream); torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1); launches = eng.kernel_launches - launches0
    eng.profile(True); eng.run(max(3, min(args.steps, 10))); eng.sync(); tm = eng.timers(); eng.profile(False)
    blocks, nodes = eng.grid_activity(); clk = eng.clock()
    print(json.dumps({"metric": METRIC, "value": units * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
                      "ms_per_step": ms / args.steps, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": f"BASELINE {args.config}: {scene.name}", "grid": [int(r) for r in scene.grid.res], "units": units,
                                 "particles": scene.particles.n if scene.particles is not None else 0,
                                 "mesh": [scene.mesh.nv, scene.mesh.nf] if scene.mesh is not None else None, "rate_floor": rate_floor_for(res)},
                      "stage_ms": {k: (v[0] / v[1] if v[1] else 0.0) for k, v in tm.items()}, "gpu_launches": int(launches), "active_nodes": nodes,
                      "sim": {"dt": clk["dt"], "t": clk["t"] + clk["inner_t"], "escaped": clk["escaped"]}, "setup_s": {"generate": t_gen}}))
    eng.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--res", type=int, default=512, help="grid resolution of the C5 dam break (512 = 64M particles)")
    ap.add_argument("--ref-res", type=int, default=64, help="grid resolution of the bounded CPU sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--quick", action="store_true", help="development: skip the e2e and cpu_baseline legs")
    ap.add_argument("--config", default="C5", choices=["C1", "C2", "C3", "C4", "C5"], help="development: another BASELINE configuration (1 GPU, short JSON); C5 = the contract workload")
    ap.add_argument("--state", default="flowing", choices=["flowing", "rest"], help="flowing: the developed state the headline is measured on (see flowing_packed); rest: F = I, v = 0")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"], help="N > 1: peer = halo / migration / max|v| stored into the neighbours' memory by the engine's own kernels (CUDA IPC over NVLink); nccl = torch.distributed send/recv driven from Python (the baseline)")
    ap.add_argument("--perturb", type=float, default=0.0, help="development: random strain scale added to the rest state (0 = the named workload)")
    ap.add_argument("--sort-bricks", type=int, default=0, help="1: brick-major particle order (aep_config.sort_bricks), 0: cell-index order")
    ap.add_argument("--pin-dt", type=float, default=0.0, help="development: pin dt (s) so that runs are comparable -- the reference rule's dt follows rounding noise at near-massless nodes; the headline uses the reference rule")
    ap.add_argument("--sort-threshold", type=float, default=None, help="development: aep_config.sort_cost_threshold of the adaptive re-sort (default: the library's)")
    ap.add_argument("--sort-every", type=int, default=0, help="physical re-sort period in substeps (aep_config.sort_every); 0 = adaptive (default)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    elif args.config != "C5":
        run_engine_config(args)
    else:
        run_engine(args)


if __name__ == "__main__":
    main()
