"""Multi-GPU slab decomposition of the MPM substep: one process per GPU, `torch.distributed` (NCCL over
NVLink / NVSwitch on GPUs, gloo on CPU for the host-logic tests) moves the bytes, the C ABI does everything else.

The reference is single-process (no MPI/NCCL anywhere), so this layer is new.  Per substep and per neighbour pair:

    forces | halo(f) | grid update | all-reduce(max |v|) | dt rule, G2P | migrate particles | re-bin, P2G | halo(m,p)

* halo = the 3 node planes around a slab boundary that both neighbours scatter into; one symmetric exchange of partial
  sums per scatter, each side adds what it receives (identical totals on both ranks, shared planes updated redundantly).
* the only global scalar is max|v_i| for the dt rule (HybridSolver.cpp:878) -> one 4-byte all-reduce(max).
* particles whose new cell left the slab travel to the neighbour as 44-float records with their global id.

`SlabSolver` is backend-agnostic: `GpuSlabBackend` drives libaep_b200.so; tests/ provides a CPU backend on the oracle so the
same driver runs under gloo with world_size 2.  `LocalSlabGroup` steps several backends of ONE process in lockstep with direct
buffer swaps (several slabs on one GPU: decomposition-invariance test without NCCL).
"""
from __future__ import annotations

import ctypes as C
import json
import os
import time
from typing import List, Optional, Sequence

import numpy as np

MIGRATE_FLOATS = 44


# ------------------------------------------------------------------------------------------------ partition
class SlabPlan:
    """Cells [bounds[r], bounds[r+1]) along `axis` belong to rank r."""

    def __init__(self, axis: int, bounds: Sequence[int]):
        self.axis = int(axis); self.bounds = [int(b) for b in bounds]
        self.world = len(self.bounds) - 1
        for r in range(self.world):
            if self.bounds[r + 1] - self.bounds[r] < 4 and self.world > 1:
                raise ValueError(f"slab {r} is {self.bounds[r + 1] - self.bounds[r]} cells wide; need >= 4 (cubic support spans 3 node planes)")

    @staticmethod
    def uniform(res_axis: int, world: int, axis: int = 1) -> "SlabPlan":
        return SlabPlan(axis, [round(r * res_axis / world) for r in range(world + 1)])

    @staticmethod
    def balanced(cells_axis: np.ndarray, res_axis: int, world: int, axis: int = 1, min_width: int = 4) -> "SlabPlan":
        """Boundaries at particle-count quantiles (dam break: particles are not uniform in space)."""
        hist = np.bincount(np.asarray(cells_axis, np.int64), minlength=res_axis).astype(np.float64)
        cum = np.concatenate([[0.0], np.cumsum(hist)])
        bounds = [0]
        for r in range(1, world):
            b = int(np.searchsorted(cum, cum[-1] * r / world))
            b = max(b, bounds[-1] + min_width); b = min(b, res_axis - min_width * (world - r))
            bounds.append(b)
        bounds.append(res_axis)
        return SlabPlan(axis, bounds)

    def slab(self, rank):
        return self.axis, self.bounds[rank], self.bounds[rank + 1]

    def owner_of_cells(self, cells_axis):
        return np.clip(np.searchsorted(np.asarray(self.bounds), cells_axis, side="right") - 1, 0, self.world - 1)


# ------------------------------------------------------------------------------------------------ GPU backend
class GpuSlabBackend:
    """One slab on one GPU through the C ABI; communication buffers are torch CUDA tensors owned here."""

    def __init__(self, engine, migrate_capacity: int = 1 << 16):
        import torch
        from . import capi
        self.torch = torch; self.capi = capi
        self.e = engine; self.L = engine.L; self.h = engine.h
        self.device = torch.device("cuda", engine.cfg.device)
        self.stream = torch.cuda.ExternalStream(engine.stream, device=self.device)
        self.halo_send = {}; self.halo_recv = {}
        for what in (0, 1):
            for side in (0, 1):
                n = C.c_int64(0)
                capi.check(self.L.aep_halo_info(self.h, what, side, C.byref(n)), self.h)
                self.halo_send[(what, side)] = torch.zeros(n.value, dtype=torch.float32, device=self.device)
                self.halo_recv[(what, side)] = torch.zeros(n.value, dtype=torch.float32, device=self.device)
        self.mig_cap = int(migrate_capacity)
        self.mig_send = [torch.zeros((self.mig_cap, MIGRATE_FLOATS), dtype=torch.float32, device=self.device) for _ in range(2)]
        self.mig_recv = [torch.zeros((self.mig_cap, MIGRATE_FLOATS), dtype=torch.float32, device=self.device) for _ in range(2)]
        self.vmax = torch.zeros(1, dtype=torch.float32, device=self.device)
        self.count_dtype = torch.int64
        # sync-free migration: k_g2p lists the leavers, counts live in device memory owned here
        self.mig_counts = torch.zeros(2, dtype=torch.int64, device=self.device)
        self.mig_counts_in = torch.zeros(2, dtype=torch.int64, device=self.device)
        self.mig_host = torch.zeros(4, dtype=torch.int64).pin_memory()
        self._ck(self.L.aep_migrate_bind(self.h, C.c_void_p(self.mig_send[0].data_ptr()), C.c_void_p(self.mig_send[1].data_ptr()), self.mig_cap,
                                         C.c_void_p(self.mig_counts.data_ptr())))

    def _ck(self, rc): self.capi.check(rc, self.h)
    # stepping
    def init_begin(self): self._ck(self.L.aep_init_begin(self.h))
    def init_volumes(self): self._ck(self.L.aep_init_volumes(self.h))
    def init_dt(self): self._ck(self.L.aep_init_dt(self.h))
    def step_forces(self): self._ck(self.L.aep_step_forces(self.h))
    def step_grid(self): self._ck(self.L.aep_step_grid(self.h))
    def step_g2p(self): self._ck(self.L.aep_step_g2p(self.h))
    def step_p2g(self): self._ck(self.L.aep_step_p2g(self.h))
    # halo
    def halo_pack(self, what, side):
        t = self.halo_send[(what, side)]
        self._ck(self.L.aep_halo_pack(self.h, what, side, C.c_void_p(t.data_ptr())))
        return t
    def halo_recv_buffer(self, what, side): return self.halo_recv[(what, side)]
    def halo_add(self, what, side, t): self._ck(self.L.aep_halo_add(self.h, what, side, C.c_void_p(t.data_ptr())))
    # vmax
    def vmax_get(self):
        self._ck(self.L.aep_vmax_get(self.h, C.c_void_p(self.vmax.data_ptr()))); return self.vmax
    def vmax_set(self, t): self._ck(self.L.aep_vmax_set(self.h, C.c_void_p(t.data_ptr())))
    # migration
    def migrate_extract(self):
        nl = C.c_int64(0); nh = C.c_int64(0)
        self._ck(self.L.aep_migrate_extract(self.h, C.c_void_p(self.mig_send[0].data_ptr()), C.c_void_p(self.mig_send[1].data_ptr()),
                                            self.mig_cap, C.byref(nl), C.byref(nh)))
        return self.mig_send[0][:nl.value], self.mig_send[1][:nh.value]
    def migrate_begin(self):
        """Gather the leavers k_g2p listed into the send buffers; returns (counts_out, counts_in) device int64[2] tensors."""
        self._ck(self.L.aep_migrate_extract_begin(self.h))
        return self.mig_counts, self.mig_counts_in
    def migrate_counts_to_host(self):
        """Enqueue the D2H copy of (out_low, out_high, in_low, in_high) and return an event to wait on later."""
        torch = self.torch
        self.mig_host.copy_(torch.cat([self.mig_counts, self.mig_counts_in]), non_blocking=True)
        ev = torch.cuda.Event(); ev.record(torch.cuda.current_stream(self.device))
        return ev
    def migrate_end(self, ev):
        ev.synchronize()
        ol, oh, il, ih = (int(v) for v in self.mig_host.tolist())
        self._ck(self.L.aep_migrate_extract_end(self.h, ol, oh))
        return (self.mig_send[0][:ol], self.mig_send[1][:oh]), (il, ih)
    def step_p2g_arrivals(self, count): self._ck(self.L.aep_step_p2g_arrivals(self.h, int(count)))
    def migrate_recv_buffer(self, side, n):
        if n > self.mig_cap:
            raise RuntimeError(f"migration receive capacity {self.mig_cap} < {n}")
        return self.mig_recv[side][:n]
    def migrate_insert(self, from_low, from_high):
        self._ck(self.L.aep_migrate_insert(self.h, C.c_void_p(from_low.data_ptr()) if from_low.numel() else None, from_low.shape[0],
                                           C.c_void_p(from_high.data_ptr()) if from_high.numel() else None, from_high.shape[0]))
    def sync(self): self.e.sync()

    def close(self):
        """Drop every tensor that was used on the engine's stream BEFORE the engine (and with it the stream) is destroyed: torch's
        pinned-memory allocator records an event on that stream when such a tensor is freed."""
        self.e.sync(); self.torch.cuda.synchronize(self.device)
        self.halo_send = self.halo_recv = self.mig_send = self.mig_recv = None
        self.vmax = self.mig_counts = self.mig_counts_in = self.mig_host = None
        self.stream = None

    def particles_local(self):
        return download_local(self.e)


def download_local(engine):
    """Particles a slab context currently holds, in its own order, with their global ids (dead slots dropped)."""
    from . import capi
    from .scenes import from_colmajor, mats_from_colmajor
    n = engine.n_particles
    ids = np.empty(n, np.int64); b = {k: np.empty(3 * n) for k in ("x", "v", "B1", "B2", "B3")}
    FE = np.empty(9 * n); FP = np.empty(9 * n); vol = np.empty(n); q = np.empty(n)
    P = lambda a: a.ctypes.data_as(capi.dp)
    capi.check(engine.L.aep_download_particles_local(engine.h, ids.ctypes.data_as(capi.i64p), P(b["x"]), P(b["v"]), P(b["B1"]), P(b["B2"]), P(b["B3"]),
                                                     P(FE), P(FP), P(vol), P(q)), engine.h)
    B = np.stack([from_colmajor(b["B1"], n), from_colmajor(b["B2"], n), from_colmajor(b["B3"], n)], axis=1)
    return dict(ids=ids, x=from_colmajor(b["x"], n), v=from_colmajor(b["v"], n), B=B, FE=mats_from_colmajor(FE, n), FP=mats_from_colmajor(FP, n), vol=vol, q=q)


# ------------------------------------------------------------------------------------------------ peer-memory exchange
class PeerSlabGroup:
    """Several slab contexts of ONE process, connected by plain device pointers and stepped in lockstep by the library
    (aep_comm_connect_local / aep_group_init / aep_group_run): the same kernels and the same flag protocol as the multi-process
    path, where the pointers are CUDA IPC mappings (connect_ranks below).  One GPU or several."""

    def __init__(self, engines: List, migrate_capacity: int = 1 << 14):
        from . import capi
        self.capi = capi; self.engs = list(engines); self.L = engines[0].L
        self.world = len(engines)
        self.arr = (C.c_void_p * self.world)(*[e.h for e in engines])
        capi.check(self.L.aep_comm_connect_local(self.arr, self.world, int(migrate_capacity)), engines[0].h)

    def _ck(self, rc):
        if rc != 0:
            msgs = [self.L.aep_last_error(e.h) for e in self.engs]
            raise self.capi.AepError(f"libaep_b200 error {rc}: " + " | ".join(m.decode() for m in msgs if m))

    def init(self): self._ck(self.L.aep_group_init(self.arr, self.world))
    def run(self, n): self._ck(self.L.aep_group_run(self.arr, self.world, int(n)))
    def substep(self): self.run(1)

    def gather_particles(self):
        parts = [download_local(e) for e in self.engs]
        ids = np.concatenate([p["ids"] for p in parts]); order = np.argsort(ids)
        out = {k: np.concatenate([p[k] for p in parts], axis=0)[order] for k in parts[0] if k != "ids"}
        out["ids"] = ids[order]
        return out


def connect_ranks(engine, rank: int, world: int, migrate_capacity: int, group=None):
    """One process per GPU: export this rank's communication block, all-gather the 256-byte blobs over torch.distributed (the only
    thing the process group is used for besides timing), connect.  From then on engine.init()/run() exchange over peer memory."""
    import torch
    import torch.distributed as dist
    from . import capi
    blob = (C.c_ubyte * capi.COMM_BLOB_BYTES)()
    capi.check(engine.L.aep_comm_export(engine.h, C.cast(blob, C.c_void_p), int(migrate_capacity)), engine.h)
    dev = torch.device("cuda", engine.cfg.device) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    mine = torch.frombuffer(bytearray(bytes(blob)), dtype=torch.uint8).to(dev)
    allb = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allb, mine, group=group)
    raw = b"".join(bytes(t.cpu().numpy().tobytes()) for t in allb)
    buf = (C.c_ubyte * len(raw)).from_buffer_copy(raw)
    capi.check(engine.L.aep_comm_connect(engine.h, int(rank), int(world), C.cast(buf, C.c_void_p)), engine.h)


# ------------------------------------------------------------------------------------------------ per-rank driver
class SlabSolver:
    """One rank of the decomposition over torch.distributed (NCCL on GPUs, gloo on CPU)."""

    def __init__(self, backend, plan: SlabPlan, rank: int, group=None):
        import torch.distributed as dist
        self.dist = dist; self.b = backend; self.plan = plan; self.rank = rank; self.group = group
        self.world = plan.world
        self.nb = {0: rank - 1 if rank > 0 else None, 1: rank + 1 if rank < self.world - 1 else None}
        self.stats = {"halo_bytes": 0, "migrated": 0}

    def _batch(self, ops):
        if not ops:
            return
        for w in self.dist.batch_isend_irecv(ops):
            w.wait()

    def halo(self, what):
        dist = self.dist; ops = []; recv = {}
        for side in (0, 1):
            nb = self.nb[side]
            if nb is None:
                continue
            send = self.b.halo_pack(what, side); recv[side] = self.b.halo_recv_buffer(what, side)
            ops.append(dist.P2POp(dist.isend, send, nb, group=self.group)); ops.append(dist.P2POp(dist.irecv, recv[side], nb, group=self.group))
            self.stats["halo_bytes"] += send.numel() * send.element_size()
        self._batch(ops)
        for side, t in recv.items():
            self.b.halo_add(what, side, t)

    def allreduce_vmax(self):
        t = self.b.vmax_get()
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)
        self.b.vmax_set(t)

    def migrate(self):
        import torch
        dist = self.dist
        out = self.b.migrate_extract()                      # (to_low, to_high) record tensors
        dev = out[0].device
        n_send = {s: torch.tensor([out[s].shape[0]], dtype=torch.int64, device=dev) for s in (0, 1)}
        n_recv = {s: torch.zeros(1, dtype=torch.int64, device=dev) for s in (0, 1)}
        ops = []
        for side in (0, 1):
            nb = self.nb[side]
            if nb is None:
                if out[side].shape[0]:
                    raise RuntimeError(f"rank {self.rank}: {out[side].shape[0]} particles left the domain through side {side}")
                continue
            ops.append(dist.P2POp(dist.isend, n_send[side], nb, group=self.group)); ops.append(dist.P2POp(dist.irecv, n_recv[side], nb, group=self.group))
        self._batch(ops)
        counts = {s: int(n_recv[s].item()) for s in (0, 1)}
        ops = []; bufs = {}
        for side in (0, 1):
            nb = self.nb[side]
            bufs[side] = self.b.migrate_recv_buffer(side, counts[side])
            if nb is None:
                continue
            if out[side].shape[0]:
                ops.append(dist.P2POp(dist.isend, out[side], nb, group=self.group))
            if counts[side]:
                ops.append(dist.P2POp(dist.irecv, bufs[side], nb, group=self.group))
        self._batch(ops)
        self.b.migrate_insert(bufs[0], bufs[1])
        self.stats["migrated"] += out[0].shape[0] + out[1].shape[0]

    def migrate_overlapped(self):
        """Migration whose host round trip hides behind the P2G of the resident particles:
        gather leavers | exchange counts (device to device) | counts to host, async | re-sort policy + P2G (resident) |
        wait for the counts | exchange records | append | P2G (arrivals only)."""
        dist = self.dist; b = self.b
        c_out, c_in = b.migrate_begin()
        ops = []
        for side in (0, 1):
            nb = self.nb[side]
            if nb is None:
                continue
            ops.append(dist.P2POp(dist.isend, c_out[side:side + 1], nb, group=self.group)); ops.append(dist.P2POp(dist.irecv, c_in[side:side + 1], nb, group=self.group))
        self._batch(ops)
        ev = b.migrate_counts_to_host()
        b.step_p2g()
        out, counts_in = b.migrate_end(ev)
        ops = []; bufs = {}
        for side in (0, 1):
            nb = self.nb[side]
            n_in = counts_in[side] if nb is not None else 0
            bufs[side] = b.migrate_recv_buffer(side, n_in)
            if nb is None:
                if out[side].shape[0]:
                    raise RuntimeError(f"rank {self.rank}: {out[side].shape[0]} particles left the domain through side {side}")
                continue
            if out[side].shape[0]:
                ops.append(dist.P2POp(dist.isend, out[side], nb, group=self.group))
            if n_in:
                ops.append(dist.P2POp(dist.irecv, bufs[side], nb, group=self.group))
        self._batch(ops)
        b.migrate_insert(bufs[0], bufs[1])
        b.step_p2g_arrivals(bufs[0].shape[0] + bufs[1].shape[0])
        self.stats["migrated"] += out[0].shape[0] + out[1].shape[0]

    def _on_stream(self):
        """torch.distributed orders its sends / receives against torch's CURRENT stream, the engine's kernels (halo pack / add,
        migration gather, max|v|) run on the context's own non-blocking stream: every exchange must be issued with that stream
        current, or halves of a halo travel before they are packed.  Entered here, so that no caller can get it wrong."""
        st = getattr(self.b, "stream", None)
        if st is None:
            import contextlib
            return contextlib.nullcontext()
        return self.b.torch.cuda.stream(st)

    def init(self):
        with self._on_stream():
            self.b.init_begin(); self.halo(0); self.b.init_volumes(); self.allreduce_vmax(); self.b.init_dt()

    def substep(self):
        with self._on_stream():
            self._substep()

    def _substep(self):
        b = self.b
        b.step_forces(); self.halo(1)
        b.step_grid(); self.allreduce_vmax()
        b.step_g2p()
        if hasattr(b, "migrate_begin"):
            self.migrate_overlapped()
        else:
            self.migrate(); b.step_p2g()
        self.halo(0)

    def run(self, n):
        for _ in range(n):
            self.substep()


# ------------------------------------------------------------------------------------------------ in-process group
class LocalSlabGroup:
    """Several slabs stepped in lockstep inside one process (all on one GPU, or on the CPU test backend): the exchange is a
    direct buffer copy.  Same backend calls, same order as SlabSolver."""

    def __init__(self, backends: List, plan: SlabPlan, overlapped: Optional[bool] = None):
        self.bs = backends; self.plan = plan; self.world = plan.world
        # overlapped: use the sync-free migration entry points (leaver lists from G2P, P2G split into resident + arrivals)
        self.overlapped = all(hasattr(b, "migrate_begin") for b in backends) if overlapped is None else overlapped

    def _fence(self):
        """Every backend enqueues on its own non-blocking stream while the buffer swaps below run on torch's current stream:
        order them with full synchronisation (this group is a single-process test vehicle, not the fast path)."""
        for b in self.bs:
            b.sync()
        dev = getattr(self.bs[0], "device", None)
        if dev is not None and getattr(dev, "type", "cpu") == "cuda":
            self.bs[0].torch.cuda.synchronize()

    def _halo(self, what):
        sends = {}
        for r, b in enumerate(self.bs):
            for side in (0, 1):
                nb = r - 1 if side == 0 else r + 1
                if 0 <= nb < self.world:
                    sends[(r, side)] = b.halo_pack(what, side)
        self._fence()
        sends = {k: t.clone() for k, t in sends.items()}
        self._fence()
        recvs = {}
        for r, b in enumerate(self.bs):
            for side in (0, 1):
                nb = r - 1 if side == 0 else r + 1
                if 0 <= nb < self.world:
                    buf = b.halo_recv_buffer(what, side); buf.copy_(sends[(nb, 1 - side)]); recvs[(r, side)] = buf
        self._fence()
        for (r, side), buf in recvs.items():
            self.bs[r].halo_add(what, side, buf)

    def _vmax(self):
        import torch
        ts = [b.vmax_get() for b in self.bs]
        self._fence()
        m = torch.stack([t.detach().cpu() for t in ts]).max(dim=0).values
        for t in ts:
            t.copy_(m.to(t.device))
        self._fence()
        for b, t in zip(self.bs, ts):
            b.vmax_set(t)

    def _migrate(self):
        outs = [b.migrate_extract() for b in self.bs]          # synchronises each context (counts come back to the host)
        self._fence()
        outs = [tuple(t.clone() for t in o) for o in outs]
        self._fence()
        for r, b in enumerate(self.bs):
            parts = []
            for side in (0, 1):
                nb = r - 1 if side == 0 else r + 1
                if 0 <= nb < self.world:
                    src = outs[nb][1 - side]
                    buf = b.migrate_recv_buffer(side, src.shape[0]); buf.copy_(src)
                else:
                    buf = b.migrate_recv_buffer(side, 0)
                parts.append(buf)
            self._fence()
            b.migrate_insert(parts[0], parts[1])

    def _migrate_overlapped(self):
        cs = [b.migrate_begin() for b in self.bs]
        self._fence()
        for r, b in enumerate(self.bs):
            for side in (0, 1):
                nb = r - 1 if side == 0 else r + 1
                if 0 <= nb < self.world:
                    cs[r][1][side] = cs[nb][0][1 - side]
                else:
                    cs[r][1][side] = 0
        self._fence()
        evs = [b.migrate_counts_to_host() for b in self.bs]
        self._fence()
        for b in self.bs: b.step_p2g()
        ends = [b.migrate_end(ev) for b, ev in zip(self.bs, evs)]
        self._fence()
        outs = [tuple(t.clone() for t in e[0]) for e in ends]
        self._fence()
        for r, b in enumerate(self.bs):
            parts = []
            for side in (0, 1):
                nb = r - 1 if side == 0 else r + 1
                if 0 <= nb < self.world:
                    src = outs[nb][1 - side]
                    assert src.shape[0] == ends[r][1][side]
                    buf = b.migrate_recv_buffer(side, src.shape[0]); buf.copy_(src)
                else:
                    assert outs[r][side].shape[0] == 0, "particles left the domain"
                    buf = b.migrate_recv_buffer(side, 0)
                parts.append(buf)
            self._fence()
            b.migrate_insert(parts[0], parts[1])
            b.step_p2g_arrivals(parts[0].shape[0] + parts[1].shape[0])

    def init(self):
        for b in self.bs: b.init_begin()
        self._halo(0)
        for b in self.bs: b.init_volumes()
        self._vmax()
        for b in self.bs: b.init_dt()

    def substep(self):
        for b in self.bs: b.step_forces()
        self._halo(1)
        for b in self.bs: b.step_grid()
        self._vmax()
        for b in self.bs: b.step_g2p()
        if self.overlapped:
            self._migrate_overlapped()
        else:
            self._migrate()
            for b in self.bs: b.step_p2g()
        self._halo(0)

    def run(self, n):
        for _ in range(n):
            self.substep()

    def gather_particles(self):
        """All particles of all slabs, ordered by global id."""
        parts = [b.particles_local() for b in self.bs]
        ids = np.concatenate([p["ids"] for p in parts]); order = np.argsort(ids)
        out = {k: np.concatenate([p[k] for p in parts], axis=0)[order] for k in parts[0] if k != "ids"}
        out["ids"] = ids[order]
        return out


def split_scene_particles(scene, plan: SlabPlan, rank: int):
    """Indices (global ids) of the particles of `scene` that start in `rank`'s slab."""
    g = scene.grid; p = scene.particles
    a = plan.axis
    cells = np.floor((p.x[:, a] - g.mn[a]) / g.h[a]).astype(np.int64)
    return np.nonzero(plan.owner_of_cells(cells) == rank)[0]


def make_gpu_slab_engine(scene, plan: SlabPlan, rank: int, device: int, capacity_factor: float = 1.3, dt_rate_floor=None, ids=None, **engine_kw):
    """Engine holding only `rank`'s particles of `scene` (global ids preserved when the slab is a contiguous id range or ids given)."""
    import copy
    from .engine import Engine
    from .scenes import Particles
    idx = split_scene_particles(scene, plan, rank) if ids is None else ids
    p = scene.particles
    local = Particles(x=p.x[idx], v=p.v[idx], B=p.B[idx], FE=p.FE[idx], FP=p.FP[idx], m=p.m[idx], vol=p.vol[idx], q=p.q[idx],
                      E=p.E, nu=p.nu, thetaC=p.thetaC, thetaS=p.thetaS)
    shell = copy.copy(scene); shell.particles = None
    cap = int(max(1024, capacity_factor * len(idx) + 4096))
    eng = Engine(shell, device=device, particle_capacity=cap, slab=plan.slab(rank), dt_rate_floor=dt_rate_floor, **engine_kw)
    return eng, local, idx


# ------------------------------------------------------------------------------------------------ bench (N > 1)
class PeerRank:
    """One rank of the peer-memory decomposition (one process per GPU): after connect_ranks the engine's own init / run exchange halo
    planes, migrating particles and max|v| through the neighbours' memory; nothing per substep happens in Python or on the host."""

    def __init__(self, engine, rank, world, migrate_capacity):
        self.e = engine; self.rank = rank; self.world = world
        connect_ranks(engine, rank, world, migrate_capacity)

    def init(self): self.e.init()
    def run(self, n): self.e.run(n)
    def sync(self): self.e.sync()
    def close(self): self.e.sync()


def bench_main(args, workload_config, ClockSampler, measured_peak_gbs, METRIC, UNIT):
    """Strong scaling of the C5 dam break over `world` GPUs (one process per GPU, launched by torchrun).  Every rank generates its
    rows of the SAME scene the one-GPU run holds (bench.dam_break_positions: one random stream per lattice row)."""
    import torch
    import torch.distributed as dist
    import bench as B
    from . import scenes as sc
    from . import capi
    from .engine import Engine
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"bench.py --gpus {args.gpus} must be launched with torchrun --nproc-per-node {args.gpus} (WORLD_SIZE={world})")
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    res = args.res
    # dam break column spans y in [2h, 1-2h] uniformly -> uniform y-slabs are balanced for all time
    lo_c, hi_c = 2, res - 2
    bounds = [0] + [lo_c + round(r * (hi_c - lo_c) / world) for r in range(1, world)] + [res]
    plan = SlabPlan(1, bounds)
    h = 1.0 / res
    t_gen = time.perf_counter()
    x, id_base = B.dam_break_positions(res, y_cells=(plan.bounds[rank], plan.bounds[rank + 1]))
    n_local = x.shape[0]
    counts = torch.zeros(world, dtype=torch.int64, device="cuda"); counts[rank] = n_local
    dist.all_reduce(counts); counts = counts.cpu().numpy(); n_total = int(counts.sum())
    assert id_base == int(counts[:rank].sum())
    mass = sc.SAND_RHO * h ** 3 / 8.0
    arrs, keep = B.packed_rest_state(x, mass, pinned=True); del x
    B.set_state(arrs, args)
    t_gen = time.perf_counter() - t_gen
    shell = B.make_shell_scene(res)
    rate_floor = B.rate_floor_for(res)
    peer = args.exchange == "peer"
    mig_cap = max(1 << 16, (n_total // world) // 20)               # the same on every rank: the peers' buffer layouts must agree

    # context + communication buffers (setup: allocation, no data), created once; every leg uploads into it
    eng = Engine(shell, device=local, particle_capacity=int(1.25 * n_local + 65536), slab=plan.slab(rank), dt_rate_floor=rate_floor, sort_every=args.sort_every,
                 sort_cost_threshold=getattr(args, "sort_threshold", None))
    capi.check(eng.L.aep_set_particle_id_base(eng.h, id_base), eng.h)
    if peer:
        solver = PeerRank(eng, rank, world, mig_cap); be = None
        stream = torch.cuda.ExternalStream(eng.stream, device=torch.device("cuda", local))
    else:
        be = GpuSlabBackend(eng, migrate_capacity=mig_cap); solver = SlabSolver(be, plan, rank); stream = be.stream

    def load():
        eng.upload_packed(n_local, arrs, sc.SAND_E, sc.SAND_NU, 2.5e-2, 7.5e-3)

    load()
    dist.barrier()                                                         # the ranks start together: a peer's flag is waited for on the device, for seconds at most
    with torch.cuda.stream(stream):
        solver.init()
        solver.run(args.warmup)
        eng.sync(); dist.barrier(); torch.cuda.synchronize()
        sampler = ClockSampler(local) if rank == 0 else None             # NVML every 5 ms: the timed region of the 8-GPU run is ~50 ms
        l0 = eng.kernel_launches; c0 = eng.counters(); k0 = eng.clock(); m0 = eng.migration()
        dist.barrier(); torch.cuda.synchronize()
        ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        solver.run(args.steps)
        ev1.record(stream); torch.cuda.synchronize()
        ms_local = ev0.elapsed_time(ev1)
        dist.barrier()
    t = torch.tensor([ms_local], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
    clocks = sampler.stop() if sampler else None
    launches = eng.kernel_launches - l0
    c1 = eng.counters(); clk = eng.clock(); m1 = eng.migration()
    mig_local = (m1["sent"] - m0["sent"]) if peer else solver.stats["migrated"]
    agg = torch.tensor([eng.grid_activity()[1], eng.n_particles, mig_local, c1["sorts"] - c0["sorts"]], dtype=torch.int64, device="cuda"); dist.all_reduce(agg)
    nodes, n_now, migrated, sorts_all = (int(v) for v in agg.tolist())
    assert n_now == n_total, f"particles were lost in migration: {n_now} of {n_total}"
    value = n_total * args.steps / (ms * 1e-3)
    sim_s = (clk["t"] + clk["inner_t"]) - (k0["t"] + k0["inner_t"])
    # per-stage time on this rank (profiled pass: every stage fenced by events, so waiting for the neighbours shows up under "halo")
    with torch.cuda.stream(stream):
        eng.profile(True); solver.run(3); eng.sync(); tm = eng.timers(); eng.profile(False)
    dist.barrier()
    stage_ms = {k: (v[0] / v[1] if v[1] else 0.0) for k, v in tm.items()}
    peak, peak_kind = measured_peak_gbs()
    n_here = eng.n_particles; nodes_here = eng.grid_activity()[1]
    dom, roofline = B.kernel_roofline(stage_ms, n_here, nodes_here, peak, peak_kind)
    sub_bytes = B.BYTES_PARTICLE[sc.SAND] * n_total + B.BYTES_NODE * nodes
    sub_gbs = sub_bytes / (ms * 1e-3 / args.steps) / 1e9 / world
    roofline["rank"] = 0
    roofline["substep"] = {"algorithmic_bytes": sub_bytes, "achieved_gbs_per_gpu": sub_gbs, "frac": sub_gbs / peak, "active_nodes": nodes}
    roofline["p2g_g2p"] = B.transfers_roofline(stage_ms, n_here, nodes_here, peak)
    plane_bytes = res * res * 16
    halo_bytes = (2 if 0 < rank < world - 1 else 1) * (4 + 3) * plane_bytes if peer else solver.stats["halo_bytes"] / max(1, args.steps + args.warmup + 3 + 1)
    # e2e: upload from pinned host memory + init + K substeps + f32 positions back, wall clock max over ranks
    out_t = torch.empty((int(1.25 * n_local + 65536), 3), dtype=torch.float32, pin_memory=True)
    eng.sync(); dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    load()
    dist.barrier()
    with torch.cuda.stream(stream):
        solver.init(); solver.run(args.steps)
    capi.check(eng.L.aep_download_positions_f32(eng.h, C.cast(out_t.data_ptr(), C.POINTER(C.c_float))), eng.h)
    torch.cuda.synchronize(); dist.barrier()
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda"); dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e = {"value": n_total * args.steps / float(t_e2e.item()), "unit": UNIT, "h2d_bytes_per_step": 36 * 8 * n_total / args.steps,
           "d2h_bytes_per_step": 12 * n_total / args.steps, "seconds": float(t_e2e.item()),
           "what": "per rank, on a context created (and sized) beforehand: aep_upload_particles(fp64 host, pinned) + init + K substeps (halo / migration / max|v| exchange included) + f32 positions"}
    eng.sync(); dist.barrier()
    if be is not None:
        be.close()
    else:
        solver.close()
    del solver, be
    if rank == 0:
        exch = ("peer memory: the engine's kernels store halo planes (4 + 3 node planes per side and substep), migrating particles and max|v| into the neighbours' "
                "memory (CUDA IPC mappings over NVLink) and publish epoch flags; one CUDA-graph launch per substep, no host synchronisation, no NCCL on the data path"
                if peer else "NCCL send/recv driven from Python: 2 halo exchanges (3 node planes per side) + 1 four-byte all-reduce(max) + particle migration per substep")
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": dict(workload_config(args, n_particles=n_total), decomposition=f"{world} y-slabs, bounds {plan.bounds}", exchange=exch),
                "roofline": roofline, "cpu_baseline": None, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
                "comm": {"halo_bytes_per_step_per_rank": halo_bytes, "migrated_particles": migrated, "migrated_particles_rank0": int(mig_local)},
                "running": {"state": args.state, "sorts_in_timed_region_all_ranks": sorts_all, "migrated_particles_in_timed_region": migrated,
                            "amortised_sort_ms": (c1["sorts"] - c0["sorts"]) * stage_ms.get("sort", 0.0) / args.steps,
                            "simulated_seconds_per_wall_second": sim_s / (ms * 1e-3) if ms > 0 else 0.0},
                "sim": {"dt": clk["dt"], "t": clk["t"] + clk["inner_t"], "escaped": clk["escaped"], "vmax": clk["vmax"]}, "setup_s": {"generate": t_gen}}
        print(json.dumps(line))
    eng.close()
    dist.destroy_process_group()
